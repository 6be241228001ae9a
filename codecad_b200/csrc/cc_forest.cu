// Union forests: dense grid_eval of a program that is one tree of (rounded) unions over fused
// primitives (BASELINE config 5: 500 rounded boxes), with exact culling.
//
// The reference evaluates every primitive at every point (codecad/nodes/codegen.py:17-63 walks the
// whole program).  Here a CTA owns a 32^3 super-tile of the grid: it decides once which primitives
// can influence the super-tile at all (a few dozen at most, out of hundreds), every warp then
// repeats the decision for each 16^3 tile against that short list, and only what is left — three
// primitives on average in config 5 — is evaluated, with the very same op library as the other
// kernels (cc_ops.cuh).  The outputs are bit-identical to the full evaluation; the argument
//  in short, for a region R (super-tile or tile):
//
//   * per primitive k the loader supplies a ball bound  lb_k <= w_k(p) <= ub_k  for the COMPUTED
//     fp32 distance of every point p of R (cc_program.cpp prim_bounds + the error slack);
//   * U = min_k ub_k bounds R's final values from above (a union never exceeds its nearer operand);
//     tau = max(U, rmax + 1.01 (rmax + D)) + slack, with rmax the largest blend radius and D the
//     largest depth of a primitive R may be inside of; primitives with lb_k > tau are "far";
//   * cc-arith's rounded union (cc_ops.cuh) returns its nearer operand UNCHANGED whenever the
//     other one is farther than r + 1.01 |r - w_near|, and is a plain minimum when both operands
//     are farther than r.  By induction over the tree every node is then either bit-identical with
//     and without the far primitives, or larger than tau in both evaluations — and the root is at
//     most U <= tau, so it is identical.  (tau of a tile is at most tau of its super-tile, so what
//     the super-tile dropped stays dropped.)
//
// What remains is a small stack program (PUSH / PRIM / COMBINE); the stack lives in shared memory,
// [level][point][thread] float4.  The first launch provides CCF_STACK_CAP levels and CCF_EVENT_CAP
// events (4 CTAs per SM); the rare super-tile that needs more is listed and redone by a second
// launch with the program's full depth.  Compile with -fmad=false (cc_math.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "cc_internal.h"
#include "cc_ops.cuh"

#define CCF_THREADS 256
#define CCF_SUPER 32      // cells per axis of a CTA's super-tile: 2 x 2 x 2 tiles of 16^3
#define CCF_TILE 16       // each tile is culled again against the super-tile's short list, by every warp for itself,
                          // and evaluated as 2 x 2 x 2 bricks of 8^3 cells (two points per thread)
#define CCF_LEAF_CAP 64   // tiles refine only if the super-tile keeps at most this many primitives (one 64-bit mask)
#define CCF_STACK_CAP 4   // stack levels / events the first launch provides; super-tiles that need more go to a
#define CCF_EVENT_CAP 256 // second launch with the program's full depth and event count (rare: dense clusters)

struct cc_forest_args {
    cc_eval_args a;
    cc_forest_launch f;
    uint32_t sx, sy, sz;       // super-tiles per axis
    uint32_t stack_cap, ev_cap;
    uint32_t region0_f4;       // float4s of the first shared region: max(stack, prefix counts)
    uint32_t *overflow;        // [0] = count, [1..] = super-tile ids deferred to the second launch
    const uint32_t *work;      // second launch: the list written by the first one (same layout); nullptr = all
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

// one point, parameters that differ from lane to lane (the masked form covers full matrices too)
template <bool RECT>
__device__ __forceinline__ void ccf_prim_lane(const uint32_t *__restrict__ code, uint32_t pc, const float (&x)[1],
                                              const float (&y)[1], const float (&z)[1], cc_val<float> (&L)[1])
{
    const float4 *q = reinterpret_cast<const float4 *>(code + pc);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5),
                 g = __ldg(q + 6);
    float m[12], mf[12];
    m[0] = a.y; m[1] = a.z; m[2] = a.w; m[3] = b.x; m[4] = b.y; m[5] = b.z;
    m[6] = b.w; m[7] = c.x; m[8] = c.y; m[9] = c.z; m[10] = c.w; m[11] = d.x;
    mf[0] = e.y; mf[1] = e.z; mf[2] = e.w; mf[3] = f.x; mf[4] = f.y; mf[5] = f.z;
    mf[6] = f.w; mf[7] = g.x; mf[8] = g.y; mf[9] = g.z; mf[10] = 0.f; mf[11] = 0.f;
    const uint32_t masks = __float_as_uint(g.w);  // word 27 is written for every fused primitive
    cc_prim_n<RECT, true, float, 1>(m, mf, masks & 0x1FFu, (masks >> 9) & 0x1FFu, d.y, d.z, d.w, e.x, x, y, z, L);
}

// lb <= w_k(p) <= ub for every point p within `rad` of centre (mx, my, mz)   (cc_program.cpp prim_bounds)
__device__ __forceinline__ void ccf_bounds(const float4 *__restrict__ bounds, uint32_t k, float mx, float my, float mz,
                                           float rad, float slack, float *lb, float *ub, float *depth)
{
    const float4 b0 = __ldg(bounds + 2 * k), b1 = __ldg(bounds + 2 * k + 1);
    const float dx = mx - b0.x, dy = my - b0.y, dz = mz - b0.z;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    *lb = b0.w * fmaxf(dist - rad, 0.0f) - b1.x - slack;
    *ub = b1.y * (dist + rad) - b1.z + slack;
    *depth = b1.w;
}

// number of set bits of mask in [a, b), 0 <= a <= b <= 64
__device__ __forceinline__ uint32_t ccf_count(unsigned long long mask, uint32_t a, uint32_t b)
{
    const unsigned long long below_b = b >= 64 ? ~0ull : ((1ull << b) - 1ull);
    const unsigned long long below_a = a >= 64 ? ~0ull : ((1ull << a) - 1ull);
    return (uint32_t)__popcll(mask & below_b & ~below_a);
}

#ifndef CCF_MIN_CTAS
#define CCF_MIN_CTAS 4  // 64 registers: four CTAs = 32 warps per SM
#endif
__global__ void __launch_bounds__(CCF_THREADS, CCF_MIN_CTAS) cc_forest_kernel(const cc_forest_args A)
{
    typedef float2 V;
    typedef cc_val<V> Val;
    extern __shared__ float4 smem4[];
    // [stack_cap][2 rows][CCF_THREADS] float4 stack (its first words double as the prefix counts of phase S)
    // | super-tile program [ev_cap] uint2 | per-warp tile programs [8][ev_cap] uint32 | kept leaves [CCF_LEAF_CAP]
    float4 *stack = smem4;
    uint32_t *s_prefix = reinterpret_cast<uint32_t *>(smem4);
    uint2 *s_ev = reinterpret_cast<uint2 *>(smem4 + A.region0_f4);
    uint32_t *s_wlist = reinterpret_cast<uint32_t *>(s_ev + A.ev_cap);
    uint32_t *s_leaf = s_wlist + (size_t)(CCF_THREADS / 32) * A.ev_cap;
    uint32_t *s_leafw = s_leaf + CCF_LEAF_CAP;  // the PRIM event word (pc, kind) of every kept leaf
    __shared__ uint32_t s_warp[CCF_THREADS / 32];
    __shared__ float s_red[2][CCF_THREADS / 32];
    __shared__ uint32_t s_info[2];  // events kept, stack depth they need

    const cc_eval_args &a = A.a;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_work = A.work ? A.work[0] : A.sx * A.sy * A.sz;
    const float4 *bounds = reinterpret_cast<const float4 *>(A.f.bounds);
    const float slack = A.f.slack;
    const uint32_t n = A.f.n_leaves;
    const float cx = a.cx, cy = a.cy, cz = a.cz;

    for (uint32_t item = blockIdx.x; item < n_work; item += gridDim.x) {
        uint32_t t = A.work ? A.work[1 + item] : item;
        const uint32_t super_id = t;
        const uint32_t tz = t % A.sz;   // z fastest, so that neighbouring CTAs write neighbouring memory
        t /= A.sz;
        const uint32_t ty = t % A.sy, tx = t / A.sy;
        const uint32_t x0 = tx * CCF_SUPER, y0 = ty * CCF_SUPER, z0 = tz * CCF_SUPER;
        __syncthreads();  // the previous item's shared state is no longer in use

        // ---- S1. bounds of every primitive over the super-tile -------------------------------------
        // centre and radius: cells x0 .. x0+31 per axis, points fma(step, index, corner)
        const float half = 0.5f * (float)(CCF_SUPER - 1);
        const float mx = cc_fma(a.step, (float)(x0 + a.x_offset) + half, cx);
        const float my = cc_fma(a.step, (float)y0 + half, cy);
        const float mz = cc_fma(a.step, (float)z0 + half, cz);
        const float rad = fabsf(a.step) * (half * 1.7320509f * 1.0001f);
        const uint32_t per = (n + CCF_THREADS - 1) / CCF_THREADS;  // contiguous chunk of leaves per thread
        const uint32_t k0 = min(tid * per, n), k1 = min(k0 + per, n);
        float umin = __int_as_float(0x7f800000), dmax = 0.0f;
        for (uint32_t k = k0; k < k1; ++k) {
            float lb, ub, depth;
            ccf_bounds(bounds, k, mx, my, mz, rad, slack, &lb, &ub, &depth);
            umin = fminf(umin, ub);
            if (lb < A.f.rmax) dmax = fmaxf(dmax, depth);
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

        // ---- S2. near flags -> prefix counts over the leaves, list of kept leaves ---------------------
        uint32_t nearbits = 0;  // per <= 32 (host)
        for (uint32_t k = k0; k < k1; ++k) {
            float lb, ub, depth;
            ccf_bounds(bounds, k, mx, my, mz, rad, slack, &lb, &ub, &depth);
            nearbits |= (lb <= tau ? 1u : 0u) << (k - k0);
        }
        uint32_t m;
        uint32_t run = ccf_block_scan(__popc(nearbits), s_warp, &m);
        for (uint32_t k = k0; k < k1; ++k) {
            s_prefix[k] = run;
            if (nearbits & (1u << (k - k0))) {
                if (run < CCF_LEAF_CAP) s_leaf[run] = k;
                ++run;
            }
        }
        if (tid == CCF_THREADS - 1) s_prefix[n] = m;
        __syncthreads();
        const bool refine = m <= CCF_LEAF_CAP;

        // ---- S3. compact the stack program; leaf ranges re-expressed in kept leaves -----------------
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
                const uint32_t pl = s_prefix[ev.y & 0xffffu], pm = s_prefix[ev.y >> 16], ph = s_prefix[ev.z];
                alive = (pm != pl) && (ph != pm);  // both operands keep at least one primitive
            }
            keep |= (alive ? 1u : 0u) << (e - e0);
        }
        uint32_t n_list;
        uint32_t at = ccf_block_scan(__popc(keep), s_warp, &n_list);
        const bool fits = n_list <= A.ev_cap;
        if (fits)
            for (uint32_t e = e0; e < e1; ++e)
                if (keep & (1u << (e - e0))) {
                    const uint4 ev = __ldg(events + e);
                    uint32_t where;
                    if ((ev.x & 3u) == CC_FOREST_PRIM) {
                        where = s_prefix[ev.y];
                        if (where < CCF_LEAF_CAP) s_leafw[where] = ev.x;
                    } else where = (s_prefix[ev.y & 0xffffu] & 0xffu) | ((s_prefix[ev.y >> 16] & 0xffu) << 8) | ((s_prefix[ev.z] & 0xffu) << 16);
                    s_ev[at++] = make_uint2(ev.x, where);
                }
        __syncthreads();
        if (tid == 0) {
            uint32_t depth = 0, deepest = 0;
            if (fits)
                for (uint32_t i = 0; i < n_list; ++i) {
                    const uint32_t type = s_ev[i].x & 3u;
                    if (type == CC_FOREST_PUSH) deepest = max(deepest, ++depth);
                    else if (type == CC_FOREST_COMBINE) --depth;
                }
            s_info[0] = n_list;
            s_info[1] = fits ? deepest : 0xffffffffu;
            if (!fits || deepest > A.stack_cap) {  // (never in the second launch: it has the program's full sizes)
                const uint32_t slot = atomicAdd(A.overflow, 1u);
                A.overflow[1 + slot] = super_id;
            }
        }
        __syncthreads();  // s_prefix (aliasing the stack) is dead from here on
        const uint32_t count_s = s_info[0];
        if (s_info[1] > A.stack_cap) continue;  // deferred (block-uniform)

        // ---- T. every warp: cull each 16^3 tile against the short list, evaluate its 8 bricks ------
        float4 *const mystack = stack + tid;
        uint32_t *const wlist = s_wlist + (size_t)warp * A.ev_cap;
        const uint32_t lz = tid & 7, ly = (tid >> 3) & 7, lx = tid >> 6;  // lx 0..3; the thread's second point is lx + 4
        float4 *out = reinterpret_cast<float4 *>(a.out);
        for (uint32_t tile = 0; tile < 8; ++tile) {
            const uint32_t tx0 = x0 + ((tile >> 2) & 1) * CCF_TILE, ty0 = y0 + ((tile >> 1) & 1) * CCF_TILE, tz0 = z0 + (tile & 1) * CCF_TILE;
            if (tx0 >= a.nx || ty0 >= a.ny || tz0 >= a.nz) continue;  // uniform
            uint32_t count = count_s;
            if (refine) {
                const float h2 = 0.5f * (float)(CCF_TILE - 1);
                const float ux = cc_fma(a.step, (float)(tx0 + a.x_offset) + h2, cx);
                const float uy = cc_fma(a.step, (float)ty0 + h2, cy);
                const float uz = cc_fma(a.step, (float)tz0 + h2, cz);
                const float r2 = fabsf(a.step) * (h2 * 1.7320509f * 1.0001f);
                // Bounds of a kept primitive over the tile from its OWN computed value at the tile centre:
                // the solid's distance is 1-Lipschitz, so |w(p) - w(centre)| <= g_hi |p - centre| (+ rounding,
                // inside the slack) — far tighter than the ball bound of the super-tile pass.  Lane i (and
                // i + 32) evaluates kept leaf i at the centre with the scalar form of the same op code.
                float lb0 = 0.f, lb1 = 0.f, u2 = __int_as_float(0x7f800000), d2 = 0.0f;
                const bool v0 = lane < m, v1 = lane + 32 < m;
                {
                    const float px[1] = {ux}, py[1] = {uy}, pz[1] = {uz};
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool v = hh ? v1 : v0;
                        if (hh && !__any_sync(0xffffffffu, v1)) break;  // uniform
                        const uint32_t i = v ? lane + 32 * hh : 0u;
                        const uint32_t lw = s_leafw[i], kind = (lw >> 2) & 63u, pc = lw >> 8;
                        const bool rect = kind == MOP_PRIM_RECT || kind == MOP_PRIM_RECT_M;
                        float wc = 0.0f;
                        // (the op library votes across the warp: both forms run converged, each lane keeps its own)
                        if (__any_sync(0xffffffffu, rect)) {
                            cc_val<float> Lc[1];
                            ccf_prim_lane<true>(a.code, pc, px, py, pz, Lc);
                            if (rect) wc = Lc[0].w;
                        }
                        if (__any_sync(0xffffffffu, !rect)) {
                            cc_val<float> Lc[1];
                            ccf_prim_lane<false>(a.code, pc, px, py, pz, Lc);
                            if (!rect) wc = Lc[0].w;
                        }
                        const float4 b1 = __ldg(bounds + 2 * s_leaf[i] + 1);
                        const float spread = b1.y * r2 + 2.0f * slack;
                        const float lb = wc - spread, ub = wc + spread;
                        if (v) {
                            u2 = fminf(u2, ub);
                            if (lb < A.f.rmax) d2 = fmaxf(d2, b1.w);
                        }
                        if (hh) lb1 = lb; else lb0 = lb;
                    }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    u2 = fminf(u2, __shfl_xor_sync(0xffffffffu, u2, d));
                    d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, d));
                }
                // (never above the super-tile's threshold, which holds for all of its points: what the
                // super-tile dropped, lb > tau, then stays dropped)
                const float tau2 = fminf(tau, fmaxf(u2, A.f.rmax + 1.01f * (A.f.rmax + d2 + slack)) + 2.0f * slack);
                const unsigned long long mask = (unsigned long long)__ballot_sync(0xffffffffu, v0 && lb0 <= tau2) |
                                                ((unsigned long long)__ballot_sync(0xffffffffu, v1 && lb1 <= tau2) << 32);
                count = 0;
                for (uint32_t eb = 0; eb < count_s; eb += 32) {
                    const uint32_t e = eb + lane;
                    bool alive = false;
                    uint32_t w0 = 0;
                    if (e < count_s) {
                        const uint2 ev = s_ev[e];
                        w0 = ev.x;
                        if ((ev.x & 3u) == CC_FOREST_PRIM) alive = (mask >> ev.y) & 1ull;
                        else alive = ccf_count(mask, ev.y & 0xffu, (ev.y >> 8) & 0xffu) != 0 &&
                                     ccf_count(mask, (ev.y >> 8) & 0xffu, (ev.y >> 16) & 0xffu) != 0;
                    }
                    const uint32_t b = __ballot_sync(0xffffffffu, alive);
                    if (alive) wlist[count + __popc(b & ((1u << lane) - 1u))] = w0;
                    count += __popc(b);
                }
                __syncwarp();
            }
            // the thread's first cell of the tile; bricks and the second point are constant offsets from it
            const uint32_t *const lst = refine ? wlist : reinterpret_cast<const uint32_t *>(s_ev);
            const uint32_t lstride = refine ? 1u : 2u;
            const bool inside = tx0 + CCF_TILE <= a.nx && ty0 + CCF_TILE <= a.ny && tz0 + CCF_TILE <= a.nz;  // uniform
            const size_t plane = (size_t)a.ny * a.nz;
            float4 *const out_t = out + ((size_t)(tx0 + lx) * a.ny + (ty0 + ly)) * a.nz + (tz0 + lz);
            for (uint32_t sb = 0; sb < 8; ++sb) {
                const uint32_t ox = ((sb >> 2) & 1) * 8, oy = ((sb >> 1) & 1) * 8, oz = (sb & 1) * 8;
                if (!inside && (tx0 + ox >= a.nx || ty0 + oy >= a.ny || tz0 + oz >= a.nz)) continue;  // uniform
                const uint32_t ix0 = tx0 + ox + lx, iy = ty0 + oy + ly, iz = tz0 + oz + lz;
                // grid_eval.cl:13,31: corner + step * convert_float(id), one FMA per axis (cc_body.cuh)
                const float gy = cc_fma(a.step, (float)iy, cy), gz = cc_fma(a.step, (float)iz, cz);
                V vx[1], vy[1], vz[1];
                vx[0] = make_float2(cc_fma(a.step, (float)(ix0 + a.x_offset), cx), cc_fma(a.step, (float)(ix0 + 4 + a.x_offset), cx));
                vy[0] = make_float2(gy, gy);
                vz[0] = make_float2(gz, gz);
                Val L[1];
                L[0] = Val{vbc<V>(0.f), vbc<V>(0.f), vbc<V>(0.f), vbc<V>(0.f)};
                uint32_t sp = 0;
                for (uint32_t i = 0; i < count; ++i) {
                    const uint32_t w0 = lst[i * lstride];
                    const uint32_t type = w0 & 3u, kind = (w0 >> 2) & 63u, pc = w0 >> 8;
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
                        if (kind == MOP_UNION_R) L[0] = cc_rounded_union(__uint_as_float(__ldg(a.code + pc + 1)), L[0], B);
                        else L[0] = cc_op_union(L[0], B);
                    }
                }
                // INDEX3: z + nz * (y + ny * x); a warp writes four 128-byte runs
                float4 *const o = out_t + (size_t)ox * plane + (size_t)oy * a.nz + oz;
                if (inside || (iy < a.ny && iz < a.nz)) {
                    if (inside || ix0 < a.nx) __stcs(o, cc_lane_get(L[0], 0));
                    if (inside || ix0 + 4 < a.nx) __stcs(o + 4 * plane, cc_lane_get(L[0], 1));
                }
            }
            __syncwarp();  // the warp's tile program is rewritten next
        }
    }
}

static size_t ccf_region0(uint32_t stack_cap, uint32_t n_leaves)
{
    return std::max((size_t)stack_cap * 2 * CCF_THREADS * sizeof(float4), (size_t)((n_leaves + 1 + 3) & ~3u) * 4);
}
static size_t ccf_smem(uint32_t stack_cap, uint32_t ev_cap, uint32_t n_leaves)
{
    return ccf_region0(stack_cap, n_leaves) + (size_t)ev_cap * sizeof(uint2) + (size_t)(CCF_THREADS / 32) * ev_cap * 4 +
           2 * CCF_LEAF_CAP * 4;
}

size_t cc_forest_smem_bytes(const cc_forest &f) { return ccf_smem(std::max(f.max_depth, 1u), f.n_events, f.n_leaves); }

uint32_t cc_forest_overflow_words(const cc_eval_args &a)
{
    const uint32_t sx = (a.nx + CCF_SUPER - 1) / CCF_SUPER, sy = (a.ny + CCF_SUPER - 1) / CCF_SUPER,
                   sz = (a.nz + CCF_SUPER - 1) / CCF_SUPER;
    return 1u + sx * sy * sz;
}

int cc_launch_forest(const cc_eval_args &a, const cc_forest_launch &f, uint32_t *d_overflow, int sm_count, void *stream)
{
    cc_forest_args A;
    A.a = a;
    A.f = f;
    A.sx = (a.nx + CCF_SUPER - 1) / CCF_SUPER;
    A.sy = (a.ny + CCF_SUPER - 1) / CCF_SUPER;
    A.sz = (a.nz + CCF_SUPER - 1) / CCF_SUPER;
    const uint64_t grid = (uint64_t)A.sx * A.sy * A.sz;
    if (grid == 0) return 0;
    if (grid >= (1ull << 31) || a.n_blocks != 1 || a.blocks || (f.n_events + CCF_THREADS - 1) / CCF_THREADS > 32 ||
        (f.n_leaves + CCF_THREADS - 1) / CCF_THREADS > 32)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_overflow, 0, 4, st);
    if (e != cudaSuccess) return (int)e;
    // first launch: small stack and event list (4 CTAs per SM); what does not fit is listed in d_overflow
    A.stack_cap = std::min<uint32_t>(CCF_STACK_CAP, std::max(f.max_depth, 1u));
    A.ev_cap = std::min<uint32_t>(CCF_EVENT_CAP, f.n_events);
    A.overflow = d_overflow;
    A.work = nullptr;
    A.region0_f4 = (uint32_t)(ccf_region0(A.stack_cap, f.n_leaves) / sizeof(float4));
    size_t smem = ccf_smem(A.stack_cap, A.ev_cap, f.n_leaves);
    e = cudaFuncSetAttribute(cc_forest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ccf_smem(std::max(f.max_depth, 1u), f.n_events, f.n_leaves));
    if (e != cudaSuccess) return (int)e;
    cc_forest_kernel<<<(uint32_t)grid, CCF_THREADS, smem, st>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (A.stack_cap == std::max(f.max_depth, 1u) && A.ev_cap == f.n_events) return 0;  // nothing can overflow
    // second launch: the deferred super-tiles with the program's full depth and event count
    A.stack_cap = std::max(f.max_depth, 1u);
    A.ev_cap = f.n_events;
    A.work = d_overflow;
    A.overflow = d_overflow;  // (not written: everything fits)
    A.region0_f4 = (uint32_t)(ccf_region0(A.stack_cap, f.n_leaves) / sizeof(float4));
    smem = ccf_smem(A.stack_cap, A.ev_cap, f.n_leaves);
    const uint32_t grid2 = (uint32_t)std::min<uint64_t>(grid, (uint64_t)std::max(sm_count, 1) * 2);
    cc_forest_kernel<<<grid2, CCF_THREADS, smem, st>>>(A);
    return (int)cudaGetLastError();
}
