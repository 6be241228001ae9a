// The part of the evaluation kernels that does not depend on HOW a point is evaluated:
// tile -> cell indices -> grid coordinates, then one of the four sinks (dense float4 store,
// PyMCubes store, ordered classification list, mass-property sums + list).  Shared by the
// microcode interpreter (cc_kernels.cu) and the NVRTC scene-specialised kernels (cc_jit.cpp),
// so that both produce bit-identical coordinates, layouts and list orders.
#ifndef CC_BODY_CUH
#define CC_BODY_CUH

#include "cc_device_types.h"
#include "cc_math.cuh"

#include "cc_scan.cuh"

// EVAL: functor  void operator()(const V (&gx)[G], const V (&gy)[G], const V (&gz)[G], cc_val<V> (&L)[G])
// with V, G = cc_pts<PTS> (cc_math.cuh): for PTS >= 2 the points j = 2g, 2g+1 are the two lanes
// of packed vector g.
// POINTS: the launch evaluates a list of points (a.points) instead of grid coordinates; a
// compile-time switch so that the grid kernels carry nothing for it (a run-time branch cost the
// 64-register specialised kernel 3.5 % more instructions).
template <int PTS, int SINK, class EVAL, bool POINTS = false>
CC_DEV void cc_kernel_body(const cc_eval_args &a, EVAL &eval)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[PTS * (CC_THREADS / 32)];
    __shared__ uint32_t s_base;

    const uint32_t tid = threadIdx.x;
    constexpr bool ORDERED = (SINK == CC_SINK_CLASSIFY || SINK == CC_SINK_MASS);
    // tile id: launch order for pure stores; ticket order where tiles wait on predecessors
    uint32_t tile = blockIdx.x;
    if (ORDERED) {
        if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        tile = s_tile;
    }
    const uint32_t block = tile / a.tiles_per_block;
    const uint32_t tile_in_block = tile - block * a.tiles_per_block;
    float cx = a.cx, cy = a.cy, cz = a.cz;
    if (a.blocks) {
        const cc_block_desc bd = a.blocks[block];
        cx = bd.cx; cy = bd.cy; cz = bd.cz;
    }
    const uint32_t cells = a.nx * a.ny * a.nz;  // <= 2^31 per launch (host checks)
    const uint32_t nyz = a.ny * a.nz;

    typedef typename cc_pts<PTS>::V V;
    constexpr int G = cc_pts<PTS>::G, NL = cc_lane<V>::N;
    float gx[PTS], gy[PTS], gz[PTS];
    uint32_t ix[PTS], iy[PTS], iz[PTS];
    bool valid[PTS];
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        uint32_t c = tile_in_block * (CC_THREADS * PTS) + j * CC_THREADS + tid;
        valid[j] = c < cells;
        c = valid[j] ? c : 0u;
        ix[j] = c / nyz;
        uint32_t r = c - ix[j] * nyz;
        iy[j] = r / a.nz;
        iz[j] = r - iy[j] * a.nz;
        // grid_eval.cl:13,31: corner + step * convert_float(id), one FMA per axis
        gx[j] = cc_fma(a.step, (float)(ix[j] + a.x_offset), cx);
        gy[j] = cc_fma(a.step, (float)iy[j], cy);
        gz[j] = cc_fma(a.step, (float)iz[j], cz);
        if (POINTS) {
            const float4 p = reinterpret_cast<const float4 *>(a.points)[c];
            gx[j] = p.x; gy[j] = p.y; gz[j] = p.z;
        }
    }

    float4 L[PTS];
    {
        V vx[G], vy[G], vz[G];
        cc_val<V> LV[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            vx[g] = cc_pack<V>(gx + g * NL);
            vy[g] = cc_pack<V>(gy + g * NL);
            vz[g] = cc_pack<V>(gz + g * NL);
        }
        eval(vx, vy, vz, LV);
#pragma unroll
        for (int j = 0; j < PTS; ++j) L[j] = cc_lane_get(LV[j / NL], j % NL);
    }

    if (SINK == CC_SINK_FLOAT4) {
        float4 *out = reinterpret_cast<float4 *>(a.out) + (size_t)block * cells;
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            uint32_t c = tile_in_block * (CC_THREADS * PTS) + j * CC_THREADS + tid;
            if (valid[j]) __stcs(out + c, L[j]);  // INDEX3: z + nz*(y + ny*x) == linear cell index
        }
    } else if (SINK == CC_SINK_PYMCUBES) {
        float *out = reinterpret_cast<float *>(a.out) + (size_t)block * cells;
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            // grid_eval.cl:18: z + (x + (ny - y - 1) * nx) * nz
            size_t idx = (size_t)iz[j] + ((size_t)ix[j] + (size_t)(a.ny - iy[j] - 1) * a.nx) * a.nz;
            if (valid[j]) __stcs(out + idx, L[j].w);
        }
    } else {
        // ---- classification ----
        const uint32_t lane = tid & 31, warp = tid >> 5;
        bool hit[PTS];
        uint32_t sum[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) sum[i] = 0;
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            const float v = L[j].w;
            if (SINK == CC_SINK_CLASSIFY) {
                hit[j] = valid[j] && (v > -a.threshold && v < a.threshold);  // subdivision.cl:25
            } else {
                const bool inside = valid[j] && (v <= -a.threshold);  // mass_properties.cl:31
                hit[j] = valid[j] && !inside && (v < a.threshold);    // mass_properties.cl:43
                if (inside) {
                    // mass_properties.cl:34-41, order xx,xy,xz,x,yy,yz,y,zz,z,n
                    const uint32_t x = ix[j], y = iy[j], z = iz[j];
                    sum[0] += x * x; sum[1] += x * y; sum[2] += x * z; sum[3] += x;
                    sum[4] += y * y; sum[5] += y * z; sum[6] += y; sum[7] += z * z;
                    sum[8] += z; sum[9] += 1u;
                }
            }
        }
        if (SINK == CC_SINK_MASS) {
            // hierarchical reduction: REDUX across the warp, one atomic per warp and moment
            uint32_t *dst = a.sums + (a.blocks ? (size_t)block * 10 : 0);
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                uint32_t w = __reduce_add_sync(0xffffffffu, sum[i]);
                if (lane == 0 && w) atomicAdd(dst + i, w);
            }
        }
        // The leaf level of mass_properties runs with threshold 0 (mass_properties.py:87-90): `inside` is
        // v <= -0 and a hit would need v < 0 on top of !inside — no cell can qualify, the list stays
        // empty, and the tile neither ranks nor waits for its predecessors.  (Warp-uniform: a launch
        // argument.)  The leaf launch is 98 % of a mass_properties call.
        if (SINK == CC_SINK_MASS && a.threshold == 0.0f) return;
        // ranks inside the tile, in cell order (j major, then warp, then lane)
        uint32_t ballot[PTS];
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            ballot[j] = __ballot_sync(0xffffffffu, hit[j]);
            if (lane == 0) s_warp[j * (CC_THREADS / 32) + warp] = __popc(ballot[j]);
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive scan of the PTS * 4 warp counts (<= 32 entries) with shuffles
            constexpr int NW = PTS * (CC_THREADS / 32);
            static_assert(NW <= 32, "the warp-count scan covers at most 32 entries");
            uint32_t v = (lane < NW) ? s_warp[lane] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane < NW) s_warp[lane] = incl - v;
            uint32_t base = cc_lookback(a.tile_status, tile, total, a.counter);
            if (lane == 0) {
                s_base = base;
                if (tile == gridDim.x - 1) *a.counter = base + total;  // final list length
            }
        }
        __syncthreads();
        const uint32_t base = s_base;
#pragma unroll
        for (int j = 0; j < PTS; ++j) {
            if (hit[j]) {
                uint32_t pos = base + s_warp[j * (CC_THREADS / 32) + warp] + __popc(ballot[j] & ((1u << lane) - 1u));
                reinterpret_cast<uchar4 *>(a.list)[pos] =
                    make_uchar4((unsigned char)ix[j], (unsigned char)iy[j], (unsigned char)iz[j], 0);
                if (a.list_block) a.list_block[pos] = block;
            }
        }
    }
}


#endif
