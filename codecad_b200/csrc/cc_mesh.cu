// Marching cubes over the leaf blocks of a subdivision, on the device.
//
// Replaces, for the mesh-export caller of the hot path, the per-block loop of
// /root/reference/codecad/rendering/mesh.py:36-74: one grid_eval_pymcubes launch + one blocking
// 8 MiB device->host copy + a CPU mcubes.marching_cubes() call per block.  Here all blocks of a
// chunk are evaluated by one PYMCUBES-sink launch into HBM, classified and triangulated by two
// kernels, and only the triangles come back.
//
// mcubes (PyMCubes 0.0.6, requirements.txt:11) is a third-party package that is not in the
// reference tree; its published algorithm (P. Bourke's polygonise conventions: corner/edge
// numbering, "inside" = value <= isovalue, vertices by linear interpolation in double precision)
// is restated here and in oracle/mc_oracle.py; the case table is generated (tools/make_mc_tables.py).
// Output order is deterministic: blocks in the given order, cells in (i, j, k) order with k
// fastest — mcubes' loop order — and the triangles of a cell in table order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cc_internal.h"
#include "cc_mc_table.h"

__constant__ unsigned char c_mc_count[256];
__constant__ unsigned char c_mc_tri[256][16];
__constant__ unsigned char c_mc_edge_corner[12][2];

__device__ __forceinline__ uint32_t cc_mc_edge_corner_d(uint32_t e, int which) { return c_mc_edge_corner[e][which]; }
// corner m -> offset along axis 0 (i), 1 (j), 2 (k): v0=(0,0,0) v1=(1,0,0) v2=(1,1,0) v3=(0,1,0), v4..v7 = +k
__device__ __forceinline__ uint32_t cc_mc_corner_d(uint32_t m, int axis)
{
    const uint32_t q = m & 3u;
    return axis == 0 ? (uint32_t)(q == 1u || q == 2u) : axis == 1 ? (uint32_t)(q >= 2u) : (m >> 2);
}

#define CC_MESH_THREADS 256

__device__ __forceinline__ uint32_t cc_mesh_case(const cc_mesh_args &a, const float *f, uint32_t i, uint32_t j, uint32_t k,
                                                 float (&v)[8])
{
    // corners v0..v7 = (i,j,k) + (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1)
    const size_t s0 = (size_t)a.d1 * a.d2, s1 = a.d2;
    const float *p = f + (size_t)i * s0 + (size_t)j * s1 + k;
    v[0] = p[0]; v[1] = p[s0]; v[2] = p[s0 + s1]; v[3] = p[s1];
    v[4] = p[1]; v[5] = p[s0 + 1]; v[6] = p[s0 + s1 + 1]; v[7] = p[s1 + 1];
    uint32_t c = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) c |= (v[m] <= 0.0f) ? (1u << m) : 0u;  // mcubes: v[m] <= isovalue, isovalue = 0
    return c;
}

// Two passes share this kernel.  COUNT writes every tile's number of triangles to tile_offsets[tile];
// cc_mesh_scan_kernel turns them into exclusive offsets (+ the total in *counter); EMIT reads its base
// from there.  (The first version ordered the tiles with a ticket and a decoupled look-back inside the
// emit pass: with 896 k tiles of which 95 % are empty, every CTA sat ~7 us at the barrier behind its
// look-back warp — 11.6 ms for csg_example at 512^3, stall_barrier 30 per issue; profiles/r1_mesh.md.)
template <bool EMIT>
__global__ void __launch_bounds__(CC_MESH_THREADS) cc_mesh_kernel(const cc_mesh_args a)
{
    __shared__ uint32_t s_warp[CC_MESH_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t block = tile / a.tiles_per_block;
    const uint32_t tile_in_block = tile - block * a.tiles_per_block;
    const uint32_t c0 = a.d0 - 1, c1 = a.d1 - 1, c2 = a.d2 - 1;
    const uint32_t cells = c0 * c1 * c2;
    const uint32_t cell = tile_in_block * CC_MESH_THREADS + tid;
    const bool valid = cell < cells;
    uint32_t i = 0, j = 0, k = 0, cs = 0, n = 0;
    float v[8];
    const float *f = a.field + (size_t)block * a.d0 * a.d1 * a.d2;
    if (valid) {
        i = cell / (c1 * c2);
        const uint32_t r = cell - i * (c1 * c2);
        j = r / c2;
        k = r - j * c2;
        cs = cc_mesh_case(a, f, i, j, k, v);
        n = c_mc_count[cs];
    }
    // triangles of this tile: warp scan + scan of the warp totals
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < CC_MESH_THREADS / 32) ? s_warp[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
        if (lane < CC_MESH_THREADS / 32) s_warp[lane] = wi - w;
        if (!EMIT && lane == 0) a.tile_offsets[tile] = total;
    }
    if (!EMIT) return;
    __syncthreads();
    if (n == 0) return;
    uint32_t pos = a.tile_offsets[tile] + s_warp[warp] + (incl - n);
    const double cx = a.corner[3 * (size_t)block + 0], cy = a.corner[3 * (size_t)block + 1], cz = a.corner[3 * (size_t)block + 2];
    const double base_idx[3] = {(double)i, (double)j, (double)k};
    for (uint32_t t = 0; t < n; ++t, ++pos) {
        double out[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint32_t e = c_mc_tri[cs][3 * t + q];
            const uint32_t ca = cc_mc_edge_corner_d(e, 0), cb = cc_mc_edge_corner_d(e, 1);
            double p[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                // mc_add_vertex: c1 + (isovalue - f1) * (c2 - c1) / (f2 - f1) along the edge's axis, in double
                const double a1 = __dadd_rn(base_idx[ax], (double)cc_mc_corner_d(ca, ax));
                const double a2 = __dadd_rn(base_idx[ax], (double)cc_mc_corner_d(cb, ax));
                double val = a1;
                if (a1 != a2) {
                    const double f1 = (double)v[ca], f2 = (double)v[cb];
                    val = __dadd_rn(a1, __ddiv_rn(__dmul_rn(__dsub_rn(0.0, f1), __dsub_rn(a2, a1)), __dsub_rn(f2, f1)));
                }
                p[ax] = val;
            }
            // rendering/mesh.py:68-71: swap columns 0 and 1, negate column 1, scale, translate (float64)
            out[q][0] = __dadd_rn(__dmul_rn(p[1], a.resolution), cx);
            out[q][1] = __dadd_rn(__dmul_rn(-p[0], a.resolution), cy);
            out[q][2] = __dadd_rn(__dmul_rn(p[2], a.resolution), cz);
        }
        // rendering/mesh.py:72: triangles[:, [0, 1]] = triangles[:, [1, 0]]
        double *dst = a.vertices + (size_t)pos * 9;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            dst[0 + ax] = out[1][ax];
            dst[3 + ax] = out[0][ax];
            dst[6 + ax] = out[2][ax];
        }
        a.tri_block[pos] = a.first_block + block;
    }
}

// In-place exclusive scan of the n tile counts (n ~ 10^6), three small launches:
//   cc_scan_local_kernel  4096 elements per CTA: exclusive scan inside the CTA, CTA total -> part[cta]
//   cc_scan_parts_kernel  one CTA: exclusive scan of part[] (chunk per thread), grand total -> *counter
//   cc_scan_add_kernel    adds part[cta] to the CTA's elements
#define CC_SCAN_PER_CTA 4096u
__global__ void __launch_bounds__(1024) cc_scan_local_kernel(uint32_t *__restrict__ v, uint32_t n, uint32_t *__restrict__ part)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t i0 = blockIdx.x * CC_SCAN_PER_CTA + 4u * t;
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (i0 + k < n) ? v[i0 + k] : 0u;
    const uint32_t mine = x[0] + x[1] + x[2] + x[3];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += up;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) part[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + (incl - mine);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) v[i0 + k] = run;
        run += x[k];
    }
}

__global__ void __launch_bounds__(1024) cc_scan_parts_kernel(uint32_t *__restrict__ v, uint32_t n, uint32_t *__restrict__ counter)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += v[i];
    s_part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t add = t >= d ? s_part[t - d] : 0u;
        __syncthreads();
        s_part[t] += add;
        __syncthreads();
    }
    uint32_t run = s_part[t] - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t c = v[i];
        v[i] = run;
        run += c;
    }
    if (t == 1023u) *counter = s_part[1023];
}

__global__ void __launch_bounds__(1024) cc_scan_add_kernel(uint32_t *__restrict__ v, uint32_t n, const uint32_t *__restrict__ part)
{
    const uint32_t base = part[blockIdx.x];
    const uint32_t i0 = blockIdx.x * CC_SCAN_PER_CTA + 4u * threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < n) v[i0 + k] += base;
}

// ---- host side ------------------------------------------------------------------------------------

int cc_mesh_upload_tables(void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyToSymbolAsync(c_mc_count, cc_mc_count, sizeof(cc_mc_count), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyToSymbolAsync(c_mc_tri, cc_mc_tri, sizeof(cc_mc_tri), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyToSymbolAsync(c_mc_edge_corner, cc_mc_edge_corner, sizeof(cc_mc_edge_corner), 0,
                                    cudaMemcpyHostToDevice, st);
    return (int)e;
}

uint32_t cc_mesh_tiles_per_block(uint32_t d0, uint32_t d1, uint32_t d2)
{
    const uint64_t cells = (uint64_t)(d0 - 1) * (d1 - 1) * (d2 - 1);
    return (uint32_t)((cells + CC_MESH_THREADS - 1) / CC_MESH_THREADS);
}

int cc_launch_mesh(const cc_mesh_args &a, bool emit, void *stream)
{
    const uint32_t grid = a.n_blocks * a.tiles_per_block;
    if (grid == 0) return 0;
    if (emit) {
        cc_mesh_kernel<true><<<grid, CC_MESH_THREADS, 0, (cudaStream_t)stream>>>(a);
    } else {
        cc_mesh_kernel<false><<<grid, CC_MESH_THREADS, 0, (cudaStream_t)stream>>>(a);
        // exclusive offsets per tile + total; the CTA partials live right behind the tile counts
        const uint32_t ctas = (grid + CC_SCAN_PER_CTA - 1) / CC_SCAN_PER_CTA;
        uint32_t *part = a.tile_offsets + grid;
        cc_scan_local_kernel<<<ctas, 1024, 0, (cudaStream_t)stream>>>(a.tile_offsets, grid, part);
        cc_scan_parts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(part, ctas, a.counter);
        cc_scan_add_kernel<<<ctas, 1024, 0, (cudaStream_t)stream>>>(a.tile_offsets, grid, part);
    }
    return (int)cudaGetLastError();
}
