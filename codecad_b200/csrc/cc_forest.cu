// Union forests: dense grid_eval of a program that is one tree of (rounded) unions over fused
// primitives (BASELINE config 5: 500 rounded boxes), with exact per-tile culling.
//
// The reference evaluates every primitive at every point (codecad/nodes/codegen.py:17-63 walks the
// whole program).  Here a CTA owns a 16 x 16 x 16 tile of the grid, first decides which primitives
// can influence the tile at all, and evaluates only those — a handful out of hundreds — with the
// very same op library as the other kernels (cc_ops.cuh).  The outputs are bit-identical to the
// full evaluation; the argument in short:
//
//   * per primitive k the loader supplies a ball bound  lb_k <= w_k(p) <= ub_k  for the COMPUTED
//     fp32 distance of every point p of the tile (cc_program.cpp prim_bounds + the error slack);
//   * U = min_k ub_k bounds the tile's final values from above (a union never exceeds its nearer
//     operand); tau = max(U, rmax + 1.01 (rmax + D)) + slack, with rmax the largest blend radius and
//     D the largest depth of a primitive the tile may be inside of; primitives with lb_k > tau are "far";
//   * cc-arith's rounded union (cc_ops.cuh) returns its nearer operand UNCHANGED whenever the
//     other one is farther than r + 1.01 |r - w_near|, and is a plain minimum when both operands
//     are farther than r.  By induction over the tree every node is then either bit-identical with
//     and without the far primitives, or larger than tau in both evaluations — and the root is at
//     most U <= tau, so it is identical.
//
// What remains after culling is evaluated as a small stack program (PUSH / PRIM / COMBINE) that all
// warps of the CTA walk together; the stack lives in shared memory, [level][point][thread] float4.
// Compile with -fmad=false (cc_math.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "cc_internal.h"
#include "cc_ops.cuh"

#define CCF_THREADS 256
#define CCF_TILE 16  // cells per axis of a CTA's tile; evaluated as 2 x 2 x 2 bricks of 8^3 (two points per thread)

struct cc_forest_args {
    cc_eval_args a;
    cc_forest_launch f;
    uint32_t tiles_x, tiles_y, tiles_z;
};

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = sum
__device__ __forceinline__ uint32_t ccf_block_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();  // s_warp may still be read from a previous scan
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < CCF_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < (int)warp) base += c;
        sum += c;
    }
    *total = sum;
    return base + incl - v;
}

template <bool RECT, bool MASKED, class V, int G>
__device__ __forceinline__ void ccf_prim(const uint32_t *__restrict__ code, uint32_t pc, const V (&x)[G], const V (&y)[G],
                                         const V (&z)[G], cc_val<V> (&L)[G])
{
    // words: 1..12 m,o | 13 a | 14 b | 15 h | 16 d | 17..25 m' | 26 scale | 27 masks   (cc_kernels.cu cc_prim)
    const float4 *q = reinterpret_cast<const float4 *>(code + pc);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5),
                 g = __ldg(q + 6);
    float m[12], mf[12];
    m[0] = a.y; m[1] = a.z; m[2] = a.w; m[3] = b.x; m[4] = b.y; m[5] = b.z;
    m[6] = b.w; m[7] = c.x; m[8] = c.y; m[9] = c.z; m[10] = c.w; m[11] = d.x;
    mf[0] = e.y; mf[1] = e.z; mf[2] = e.w; mf[3] = f.x; mf[4] = f.y; mf[5] = f.z;
    mf[6] = f.w; mf[7] = g.x; mf[8] = g.y; mf[9] = g.z; mf[10] = 0.f; mf[11] = 0.f;
    const uint32_t masks = MASKED ? __float_as_uint(g.w) : 0u;
    cc_prim_n<RECT, MASKED, V, G>(m, mf, masks & 0x1FFu, (masks >> 9) & 0x1FFu, d.y, d.z, d.w, e.x, x, y, z, L);
}

__global__ void __launch_bounds__(CCF_THREADS) cc_forest_kernel(const cc_forest_args A)
{
    typedef float2 V;
    typedef cc_val<V> Val;
    extern __shared__ float4 smem4[];
    // [max_depth][2 rows][CCF_THREADS] float4 stack | prefix counts [n_leaves + 1] | compacted program [n_events] uint2
    float4 *stack = smem4;
    uint32_t *s_prefix = reinterpret_cast<uint32_t *>(smem4 + (size_t)A.f.max_depth * 2 * CCF_THREADS);
    uint2 *s_list = reinterpret_cast<uint2 *>(s_prefix + ((A.f.n_leaves + 1 + 3) & ~3u));
    __shared__ uint32_t s_warp[CCF_THREADS / 32];
    __shared__ float s_red[2][CCF_THREADS / 32];
    __shared__ uint32_t s_count;

    const cc_eval_args &a = A.a;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile -> block of the launch, tile coordinates (z fastest, so that neighbouring CTAs write neighbouring memory)
    const uint32_t tiles_per_block = A.tiles_x * A.tiles_y * A.tiles_z;
    const uint32_t block = blockIdx.x / tiles_per_block;
    uint32_t t = blockIdx.x - block * tiles_per_block;
    const uint32_t tz = t % A.tiles_z;
    t /= A.tiles_z;
    const uint32_t ty = t % A.tiles_y, tx = t / A.tiles_y;
    float cx = a.cx, cy = a.cy, cz = a.cz;
    if (a.blocks) {
        const cc_block_desc bd = a.blocks[block];
        cx = bd.cx; cy = bd.cy; cz = bd.cz;
    }
    const uint32_t x0 = tx * CCF_TILE, y0 = ty * CCF_TILE, z0 = tz * CCF_TILE;

    // ---- 1. bounds of every primitive over the tile --------------------------------------------
    // tile centre and radius: cells x0 .. x0+15 per axis, points fma(step, index, corner)
    const float half = 0.5f * (float)(CCF_TILE - 1);
    const float mx = cc_fma(a.step, (float)(x0 + a.x_offset) + half, cx);
    const float my = cc_fma(a.step, (float)y0 + half, cy);
    const float mz = cc_fma(a.step, (float)z0 + half, cz);
    const float rad = fabsf(a.step) * (half * 1.7320509f * 1.0001f);
    const float slack = A.f.slack;
    const uint32_t n = A.f.n_leaves;
    const uint32_t per = (n + CCF_THREADS - 1) / CCF_THREADS;  // contiguous chunk of leaves per thread
    const uint32_t k0 = min(tid * per, n), k1 = min(k0 + per, n);
    const float4 *bounds = reinterpret_cast<const float4 *>(A.f.bounds);
    float umin = __int_as_float(0x7f800000), dmax = 0.0f;
    for (uint32_t k = k0; k < k1; ++k) {
        const float4 b0 = __ldg(bounds + 2 * k), b1 = __ldg(bounds + 2 * k + 1);
        const float dx = mx - b0.x, dy = my - b0.y, dz = mz - b0.z;
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float lb = b0.w * fmaxf(dist - rad, 0.0f) - b1.x - slack;
        const float ub = b1.y * (dist + rad) - b1.z + slack;
        umin = fminf(umin, ub);
        if (lb < A.f.rmax) dmax = fmaxf(dmax, b1.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, d));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, d));
    }
    if (lane == 0) {
        s_red[0][warp] = umin;
        s_red[1][warp] = dmax;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < CCF_THREADS / 32; ++w) {
        umin = fminf(umin, s_red[0][w]);
        dmax = fmaxf(dmax, s_red[1][w]);
    }
    const float tau = fmaxf(umin, A.f.rmax + 1.01f * (A.f.rmax + dmax + slack)) + 2.0f * slack;

    // ---- 2. near flags -> prefix counts over the leaves ----------------------------------------
    uint32_t cnt = 0;
    for (uint32_t k = k0; k < k1; ++k) {
        const float4 b0 = __ldg(bounds + 2 * k), b1 = __ldg(bounds + 2 * k + 1);
        const float dx = mx - b0.x, dy = my - b0.y, dz = mz - b0.z;
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float lb = b0.w * fmaxf(dist - rad, 0.0f) - b1.x - slack;
        cnt += (lb <= tau) ? 1u : 0u;
    }
    uint32_t total;
    uint32_t run = ccf_block_scan(cnt, s_warp, &total);
    for (uint32_t k = k0; k < k1; ++k) {
        const float4 b0 = __ldg(bounds + 2 * k), b1 = __ldg(bounds + 2 * k + 1);
        const float dx = mx - b0.x, dy = my - b0.y, dz = mz - b0.z;
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float lb = b0.w * fmaxf(dist - rad, 0.0f) - b1.x - slack;
        s_prefix[k] = run;
        run += (lb <= tau) ? 1u : 0u;
    }
    if (tid == CCF_THREADS - 1) s_prefix[n] = total;
    __syncthreads();

    // ---- 3. compact the stack program ----------------------------------------------------------
    const uint32_t ne = A.f.n_events;
    const uint32_t eper = (ne + CCF_THREADS - 1) / CCF_THREADS;
    const uint32_t e0 = min(tid * eper, ne), e1 = min(e0 + eper, ne);
    const uint4 *events = reinterpret_cast<const uint4 *>(A.f.events);
    uint32_t keep = 0;  // bit i: event e0 + i survives (eper <= 32 is guaranteed by the host)
    for (uint32_t e = e0; e < e1; ++e) {
        const uint4 ev = __ldg(events + e);
        bool alive;
        if ((ev.x & 3u) == CC_FOREST_PRIM) {
            alive = s_prefix[ev.y + 1] != s_prefix[ev.y];
        } else {
            const uint32_t lo = ev.y & 0xffffu, mid = ev.y >> 16, hi = ev.z;
            const uint32_t pl = s_prefix[lo], pm = s_prefix[mid], ph = s_prefix[hi];
            alive = (pm != pl) && (ph != pm);  // both operands keep at least one primitive
        }
        keep |= (alive ? 1u : 0u) << (e - e0);
    }
    uint32_t n_list;
    uint32_t at = ccf_block_scan(__popc(keep), s_warp, &n_list);
    for (uint32_t e = e0; e < e1; ++e)
        if (keep & (1u << (e - e0))) {
            const uint4 ev = __ldg(events + e);
            s_list[at++] = make_uint2(ev.x, ev.w);
        }
    if (tid == 0) s_count = n_list;
    __syncthreads();
    const uint32_t count = s_count;

    // ---- 4. evaluate the 2 x 2 x 2 bricks of the tile with the compacted program ----------------
    float4 *const mystack = stack + tid;
    const uint32_t lz = tid & 7, ly = (tid >> 3) & 7, lx = tid >> 6;  // lx 0..3; the thread's second point is lx + 4
    const uint32_t cells = a.nx * a.ny * a.nz;
    float4 *out = reinterpret_cast<float4 *>(a.out) + (size_t)block * cells;
    for (uint32_t sb = 0; sb < 8; ++sb) {
        const uint32_t bx = x0 + ((sb >> 2) & 1) * 8, by = y0 + ((sb >> 1) & 1) * 8, bz = z0 + (sb & 1) * 8;
        if (bx >= a.nx || by >= a.ny || bz >= a.nz) continue;  // warp-uniform (block-uniform)
        const uint32_t ix0 = bx + lx, ix1 = bx + lx + 4, iy = by + ly, iz = bz + lz;
        // grid_eval.cl:13,31: corner + step * convert_float(id), one FMA per axis (cc_body.cuh)
        const float gy = cc_fma(a.step, (float)iy, cy), gz = cc_fma(a.step, (float)iz, cz);
        V vx[1], vy[1], vz[1];
        vx[0] = make_float2(cc_fma(a.step, (float)(ix0 + a.x_offset), cx), cc_fma(a.step, (float)(ix1 + a.x_offset), cx));
        vy[0] = make_float2(gy, gy);
        vz[0] = make_float2(gz, gz);
        Val L[1];
        L[0] = Val{vbc<V>(0.f), vbc<V>(0.f), vbc<V>(0.f), vbc<V>(0.f)};
        uint32_t sp = 0;
        for (uint32_t i = 0; i < count; ++i) {
            const uint2 ev = s_list[i];
            const uint32_t type = ev.x & 3u, kind = (ev.x >> 2) & 63u, pc = ev.x >> 8;
            if (type == CC_FOREST_PRIM) {
                switch (kind) {
                case MOP_PRIM_RECT: ccf_prim<true, false, V, 1>(a.code, pc, vx, vy, vz, L); break;
                case MOP_PRIM_RECT_M: ccf_prim<true, true, V, 1>(a.code, pc, vx, vy, vz, L); break;
                case MOP_PRIM_CIRCLE: ccf_prim<false, false, V, 1>(a.code, pc, vx, vy, vz, L); break;
                default: ccf_prim<false, true, V, 1>(a.code, pc, vx, vy, vz, L); break;
                }
            } else if (type == CC_FOREST_PUSH) {
                float4 *p = mystack + (size_t)sp * 2 * CCF_THREADS;
                p[0] = make_float4(L[0].x.x, L[0].x.y, L[0].y.x, L[0].y.y);
                p[CCF_THREADS] = make_float4(L[0].z.x, L[0].z.y, L[0].w.x, L[0].w.y);
                ++sp;
            } else {
                --sp;
                const float4 *p = mystack + (size_t)sp * 2 * CCF_THREADS;
                const float4 u = p[0], v = p[CCF_THREADS];
                const Val B{make_float2(u.x, u.y), make_float2(u.z, u.w), make_float2(v.x, v.y), make_float2(v.z, v.w)};
                // the microcode's operand order: lastValue first, the stored operand second
                if (kind == MOP_UNION_R) L[0] = cc_rounded_union(__uint_as_float(ev.y), L[0], B);
                else L[0] = cc_op_union(L[0], B);
            }
        }
        if (iy < a.ny && iz < a.nz) {
            // INDEX3: z + nz * (y + ny * x); a warp writes four 128-byte runs
            if (ix0 < a.nx) __stcs(out + ((size_t)ix0 * a.ny + iy) * a.nz + iz, cc_lane_get(L[0], 0));
            if (ix1 < a.nx) __stcs(out + ((size_t)ix1 * a.ny + iy) * a.nz + iz, cc_lane_get(L[0], 1));
        }
    }
}

size_t cc_forest_smem_bytes(const cc_forest &f)
{
    return (size_t)f.max_depth * 2 * CCF_THREADS * sizeof(float4) + (size_t)((f.n_leaves + 1 + 3) & ~3u) * 4 +
           (size_t)f.n_events * sizeof(uint2);
}

int cc_launch_forest(const cc_eval_args &a, const cc_forest_launch &f, void *stream)
{
    cc_forest_args A;
    A.a = a;
    A.f = f;
    A.tiles_x = (a.nx + CCF_TILE - 1) / CCF_TILE;
    A.tiles_y = (a.ny + CCF_TILE - 1) / CCF_TILE;
    A.tiles_z = (a.nz + CCF_TILE - 1) / CCF_TILE;
    const uint64_t grid = (uint64_t)A.tiles_x * A.tiles_y * A.tiles_z * a.n_blocks;
    if (grid == 0) return 0;
    if (grid >= (1ull << 31) || (f.n_events + CCF_THREADS - 1) / CCF_THREADS > 32) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)f.max_depth * 2 * CCF_THREADS * sizeof(float4) + (size_t)((f.n_leaves + 1 + 3) & ~3u) * 4 +
                        (size_t)f.n_events * sizeof(uint2);
    cudaError_t e = cudaFuncSetAttribute(cc_forest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cc_forest_kernel<<<(uint32_t)grid, CCF_THREADS, smem, (cudaStream_t)stream>>>(A);
    return (int)cudaGetLastError();
}
