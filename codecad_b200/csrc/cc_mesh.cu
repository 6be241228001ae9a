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
    // COUNT runs over every tile; EMIT only over the tiles that hold triangles (95 % do not), whose
    // ids the scan compacted into tile_list in increasing order
    const uint32_t tile = EMIT ? a.tile_list[blockIdx.x] : blockIdx.x;
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
    // Emission is spread over the CTA: triangle t of the tile goes to thread t mod 256, whichever
    // cell it belongs to (a serial loop per cell left the few surface threads of a warp working
    // alone on up to five triangles each, and wrote 72-byte pieces far apart).  The owning cell is
    // found by binary search in the tile's inclusive prefix sums; its corner values come from
    // shared memory.  Same arithmetic per triangle, same output order.
    __shared__ uint32_t s_incl[CC_MESH_THREADS];
    __shared__ unsigned char s_case[CC_MESH_THREADS];
    __shared__ float s_v[CC_MESH_THREADS][9];  // padded: 8 corner values per cell
    __syncthreads();
    s_incl[tid] = s_warp[warp] + incl;
    s_case[tid] = (unsigned char)cs;
    if (n) {
#pragma unroll
        for (int m = 0; m < 8; ++m) s_v[tid][m] = v[m];
    }
    __syncthreads();
    const uint32_t tile_total = s_incl[CC_MESH_THREADS - 1];
    const uint32_t tile_base = a.tile_offsets[tile];
    const double cx = a.corner[3 * (size_t)block + 0], cy = a.corner[3 * (size_t)block + 1], cz = a.corner[3 * (size_t)block + 2];
    for (uint32_t t = tid; t < tile_total; t += CC_MESH_THREADS) {
        // owner = first thread whose inclusive sum exceeds t
        uint32_t lo = 0, hi = CC_MESH_THREADS - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_incl[mid] > t) hi = mid; else lo = mid + 1;
        }
        const uint32_t owner = lo, ocs = s_case[owner];
        const uint32_t tri = t - (s_incl[owner] - c_mc_count[ocs]);  // index among the owner's triangles
        const uint32_t ocell = tile_in_block * CC_MESH_THREADS + owner;
        const uint32_t oi = ocell / (c1 * c2), orem = ocell - oi * (c1 * c2), oj = orem / c2, ok = orem - oj * c2;
        const double base_idx[3] = {(double)oi, (double)oj, (double)ok};
        const float *ov = s_v[owner];
        double out[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint32_t e = c_mc_tri[ocs][3 * tri + q];
            const uint32_t ca = cc_mc_edge_corner_d(e, 0), cb = cc_mc_edge_corner_d(e, 1);
            double p[3];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                // mc_add_vertex: c1 + (isovalue - f1) * (c2 - c1) / (f2 - f1) along the edge's axis, in double
                const double a1 = __dadd_rn(base_idx[ax], (double)cc_mc_corner_d(ca, ax));
                const double a2 = __dadd_rn(base_idx[ax], (double)cc_mc_corner_d(cb, ax));
                double val = a1;
                if (a1 != a2) {
                    const double f1 = (double)ov[ca], f2 = (double)ov[cb];
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
        const uint32_t pos = tile_base + t;
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

// In-place exclusive scan of the n tile counts (n ~ 10^6) and compaction of the ids of the non-empty
// tiles, three small launches:
//   cc_scan_local_kernel  4096 elements per CTA: exclusive scans of (count, count != 0) inside the CTA,
//                         CTA totals -> part[cta], part_f[cta]
//   cc_scan_parts_kernel  one CTA: exclusive scans of part[] and part_f[], grand totals -> counter[0..1]
//   cc_scan_add_kernel    adds part[cta] to the CTA's offsets and writes the non-empty tile ids
#define CC_SCAN_PER_CTA 4096u
__device__ __forceinline__ uint32_t cc_block_exclusive(uint32_t mine, uint32_t *s_warp, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    const uint32_t excl = s_warp[warp] + (incl - mine);
    *total = s_warp[32];
    __syncthreads();  // s_warp is reused by the next call
    return excl;
}

__global__ void __launch_bounds__(1024) cc_scan_local_kernel(uint32_t *__restrict__ v, uint32_t n, uint32_t *__restrict__ part,
                                                             uint32_t *__restrict__ flag_off, uint32_t *__restrict__ part_f)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t t = threadIdx.x;
    const uint32_t i0 = blockIdx.x * CC_SCAN_PER_CTA + 4u * t;
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (i0 + k < n) ? v[i0 + k] : 0u;
    uint32_t total, total_f;
    uint32_t run = cc_block_exclusive(x[0] + x[1] + x[2] + x[3], s_warp, &total);
    uint32_t run_f = cc_block_exclusive((x[0] != 0) + (x[1] != 0) + (x[2] != 0) + (x[3] != 0), s_warp, &total_f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) {
            v[i0 + k] = run;
            flag_off[i0 + k] = run_f | (x[k] ? 0x80000000u : 0u);  // top bit: this tile holds triangles
        }
        run += x[k];
        run_f += x[k] != 0;
    }
    if (t == 0) {
        part[blockIdx.x] = total;
        part_f[blockIdx.x] = total_f;
    }
}

// exclusive scan of v[0..n) by one CTA (chunk per thread); returns the total through *out
__device__ __forceinline__ void cc_scan_single_cta(uint32_t *__restrict__ v, uint32_t n, uint32_t *s_part, uint32_t *out)
{
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
    if (t == 1023u) *out = s_part[1023];
    __syncthreads();
}

__global__ void __launch_bounds__(1024) cc_scan_parts_kernel(uint32_t *__restrict__ part, uint32_t *__restrict__ part_f, uint32_t n,
                                                             uint32_t *__restrict__ counter)
{
    __shared__ uint32_t s_part[1024];
    cc_scan_single_cta(part, n, s_part, counter);        // counter[0] = triangles
    cc_scan_single_cta(part_f, n, s_part, counter + 1);  // counter[1] = non-empty tiles
}

__global__ void __launch_bounds__(1024) cc_scan_add_kernel(uint32_t *__restrict__ v, uint32_t n, const uint32_t *__restrict__ part,
                                                           const uint32_t *__restrict__ flag_off,
                                                           const uint32_t *__restrict__ part_f, uint32_t *__restrict__ tile_list)
{
    const uint32_t base = part[blockIdx.x], base_f = part_f[blockIdx.x];
    const uint32_t i0 = blockIdx.x * CC_SCAN_PER_CTA + 4u * threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < n) {
            v[i0 + k] += base;
            const uint32_t f = flag_off[i0 + k];
            if (f & 0x80000000u) tile_list[base_f + (f & 0x7FFFFFFFu)] = i0 + k;
        }
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

size_t cc_mesh_scratch_words(uint32_t tiles)
{
    const size_t ctas = (tiles + CC_SCAN_PER_CTA - 1) / CC_SCAN_PER_CTA;
    return 3 * (size_t)tiles + 2 * ctas;  // offsets, flag offsets, tile list; two sets of CTA partials
}

// count pass: all tiles, then the scan (counter[0] = triangles, counter[1] = non-empty tiles).
// emit pass: `emit_tiles` CTAs, one per non-empty tile.
int cc_launch_mesh(const cc_mesh_args &a, bool emit, uint32_t emit_tiles, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t grid = a.n_blocks * a.tiles_per_block;
    if (grid == 0) return 0;
    if (emit) {
        if (emit_tiles) cc_mesh_kernel<true><<<emit_tiles, CC_MESH_THREADS, 0, st>>>(a);
    } else {
        cc_mesh_kernel<false><<<grid, CC_MESH_THREADS, 0, st>>>(a);
        const uint32_t ctas = (grid + CC_SCAN_PER_CTA - 1) / CC_SCAN_PER_CTA;
        uint32_t *part = a.tile_offsets + grid, *flag_off = part + ctas, *part_f = flag_off + grid;
        cc_scan_local_kernel<<<ctas, 1024, 0, st>>>(a.tile_offsets, grid, part, flag_off, part_f);
        cc_scan_parts_kernel<<<1, 1024, 0, st>>>(part, part_f, ctas, a.counter);
        cc_scan_add_kernel<<<ctas, 1024, 0, st>>>(a.tile_offsets, grid, part, flag_off, part_f, a.tile_list);
    }
    return (int)cudaGetLastError();
}
