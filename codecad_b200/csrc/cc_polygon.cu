// 2-D outline extraction on the device: the kernel behind
// rendering.polygon2d.polygon().
//
//   cc_process_polygon_kernel   process_polygon   /root/reference/codecad/rendering/polygon2d.cl:82-175
//       (helpers encode_index :5-35, place_vertex :37-80)
//
// The grid of float4 (gradient, distance) samples that grid_eval wrote for a block is cut into
// two triangles per cell; every triangle the surface crosses gets one outline vertex (fixed-count
// gradient search on the three corner planes) and a link to the triangle that follows it along
// the outline, encoded with the side of the block it leaves through when that triangle lies in a
// neighbouring block.  Triangles whose predecessor lies outside are the starts of open chains.
//
// Differences from the OpenCL kernel, both invisible to the host algorithm: the start list is
// produced in increasing cell-index order (one tiny sort after the append; the reference's
// atomic_inc order is arbitrary), and any number of equally sized blocks is processed by one
// launch (block b reads corners + b * (cx+1)(cy+1), writes vertices/links + b * 2 cx cy and
// starts + b * max_starts).  HBM-bound: 3 float4 corner reads (L1/L2-shared between neighbours)
// + 12 bytes written per triangle.
//
// Arithmetic: -fmad=false, single IEEE operations in source order (div.rn), like the CPU oracle.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cc_internal.h"

namespace {

__device__ __forceinline__ uint32_t encode_index(int cx, int cy, uint32_t index, int gs0, int gs1)
{
    index &= (1u << 20) - 1u;  // an out-of-block index may exceed the 20-bit field
    bool is_y;
    int out_c, other_c;
    if (cx < 0 || cx >= gs0) { is_y = false; out_c = cx; other_c = cy; }
    else if (cy < 0 || cy >= gs1) { is_y = true; out_c = cy; other_c = cx; }
    else return index;
    return 0x80000000u | (is_y ? 0x40000000u : 0u) | (out_c < 0 ? 0x20000000u : 0u) | ((uint32_t)other_c << 20) | index;
}

__device__ __forceinline__ float2 place_vertex(const float2 (&pos)[3], const float4 (&val)[3])
{
    float ax = 0.f, ay = 0.f, weight = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float w = __fdiv_rn(1.f, 1.f + fabsf(val[i].w));
        ax += pos[i].x * w;
        ay += pos[i].y * w;
        weight += w;
    }
    float px = __fdiv_rn(ax, weight), py = __fdiv_rn(ay, weight);
    for (int it = 0; it < 8; ++it) {
        float gx = 0.f, gy = 0.f, residual = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float nx = val[j].x, ny = val[j].y;
            const float tmp = (nx * (px - pos[j].x) + ny * (py - pos[j].y)) + val[j].w;
            residual += tmp * tmp;
            gx += nx * tmp;
            gy += ny * tmp;
        }
        if (residual < 1e-3f) break;
        const float gl = gx * gx + gy * gy;
        if (gl < 1e-8f) break;
        const float k = __fdiv_rn(residual, gl);
        px -= gx * k;
        py -= gy * k;
    }
    return make_float2(px, py);
}

__global__ void __launch_bounds__(256) cc_process_polygon_kernel(cc_polygon_args a)
{
    const uint32_t per_block = 2u * a.cx * a.cy;
    const uint64_t gid = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (gid >= (uint64_t)per_block * a.n_blocks) return;
    const uint32_t block = (uint32_t)(gid / per_block);
    const uint32_t index = (uint32_t)(gid - (uint64_t)block * per_block);  // INDEX3_GG: t + 2 (y + cy x)
    const int t = (int)(index & 1u);
    const int y = (int)((index >> 1) % a.cy), x = (int)((index >> 1) / a.cy);
    const int gs0 = (int)a.cx, gs1 = (int)a.cy;
    const float4 *corners = a.corners + (size_t)block * (a.cx + 1) * (a.cy + 1);
    uint32_t *links = a.links + (size_t)block * per_block;

    const int off[3][2] = {{0, 0}, {1, 1}, {t, 1 - t}};
    float4 val[3];
    unsigned cell_type = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        val[i] = corners[(size_t)(y + off[i][1]) + (size_t)(gs1 + 1) * (size_t)(x + off[i][0])];
        cell_type = cell_type << 1 | (val[i].w <= 0 ? 1u : 0u);
    }
    if (cell_type == 0 || cell_type == 7) {
        links[index] = 0xFFFFFFFFu;  // entirely inside or outside
        return;
    }
    bool backwards = cell_type == 3 || cell_type == 5 || cell_type == 6;
    if (backwards) cell_type = 7 - cell_type;
    const bool flip = t == 1;
    if (flip) backwards = !backwards;
    int fx = 0, fy = 0, rx = 0, ry = 0;
    switch (cell_type) {
    case 1: fx = 0; fy = 1; rx = -1; ry = 0; break;
    case 2: fx = 0; fy = 0; rx = 0; ry = 1; break;
    default: fx = -1; fy = 0; rx = 0; ry = 0; break;  // 4
    }
    if (backwards) { int p = fx, q = fy; fx = rx; fy = ry; rx = p; ry = q; }
    if (flip) { int p = fx; fx = fy; fy = p; p = rx; rx = ry; ry = p; }
    fx += x; fy += y; rx += x; ry += y;
    links[index] = encode_index(fx, fy, (uint32_t)(1 - t) + 2u * ((uint32_t)fy + (uint32_t)gs1 * (uint32_t)fx), gs0, gs1);
    const uint32_t start_index = encode_index(rx, ry, index, gs0, gs1);
    if (start_index & 0x80000000u) {
        const uint32_t slot = atomicAdd(a.start_counter + block, 1u);
        if (slot < a.max_starts) a.starts[(size_t)block * a.max_starts + slot] = start_index ^ 0x20000000u;
    }
    float bx = a.corner_x, by = a.corner_y;
    if (a.block_corners) { bx = a.block_corners[2 * block]; by = a.block_corners[2 * block + 1]; }
    float2 pos[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        pos[i].x = bx + (float)(x + off[i][0]) * a.step;
        pos[i].y = by + (float)(y + off[i][1]) * a.step;
    }
    a.vertices[(size_t)block * per_block + index] = place_vertex(pos, val);
}

// one CTA per block: order the (few) starts by their cell index, the low 20 bits
__global__ void __launch_bounds__(1024) cc_sort_starts_kernel(uint32_t *starts, const uint32_t *counter, uint32_t max_starts)
{
    __shared__ uint32_t s[1024];
    uint32_t *mine = starts + (size_t)blockIdx.x * max_starts;
    const uint32_t n = min(counter[blockIdx.x], max_starts);
    if (n <= 1) return;  // uniform: most boxes have no open chain at all
    uint32_t m = 2;      // bitonic network over the next power of two only (typically 2-4 starts)
    while (m < n) m <<= 1;
    const uint32_t i = threadIdx.x;
    s[i] = i < n ? mine[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t p = i ^ j;
            if (p > i) {
                const uint32_t u = s[i], v = s[p];
                const uint32_t ku = u == 0xFFFFFFFFu ? 0xFFFFFFFFu : (u & 0xFFFFFu);
                const uint32_t kv = v == 0xFFFFFFFFu ? 0xFFFFFFFFu : (v & 0xFFFFFu);
                const bool up = (i & k) == 0;
                if ((ku > kv) == up) { s[i] = v; s[p] = u; }
            }
            __syncthreads();
        }
    if (i < n) mine[i] = s[i];
}

// matplotlib_slice  /root/reference/codecad/rendering/matplotlib_slice.cl:1-20: the viewer's layout
// (distance, gradient x, gradient y) at (x + y*w)*3, from the float4 grid [w][h] grid_eval wrote
__global__ void __launch_bounds__(256) cc_slice_repack_kernel(const float4 *__restrict__ field, uint32_t w, uint32_t h,
                                                             float *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;  // output pixel, x fastest
    if (i >= (uint64_t)w * h) return;
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (uint64_t)y * w);
    const float4 v = field[(size_t)y + (size_t)h * x];
    out[3 * i + 0] = v.w;
    out[3 * i + 1] = v.x;
    out[3 * i + 2] = v.y;
}

}  // namespace

int cc_launch_slice_repack(const void *d_field, uint32_t w, uint32_t h, float *d_out, void *stream)
{
    const uint64_t n = (uint64_t)w * h;
    if (n == 0) return 0;
    cc_slice_repack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4 *)d_field, w, h, d_out);
    return (int)cudaGetLastError();
}

int cc_launch_process_polygon(const cc_polygon_args &a, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t total = 2ull * a.cx * a.cy * a.n_blocks;
    if (total == 0) return 0;
    cc_process_polygon_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    cc_sort_starts_kernel<<<a.n_blocks, 1024, 0, st>>>(a.starts, a.start_counter, a.max_starts);
    return (int)cudaGetLastError();
}
