// SDF op library, device side: one __device__ function per node type of
// /root/reference/codecad/shapes/*.cl, evaluated in the canonical arithmetic of
// cc_math.cuh.  Values follow the reference convention: .xyz = unit direction
// (gradient) or point coordinates, .w = signed distance (negative inside).
//
// Each function handles ONE point; the interpreter calls it in an unrolled loop over the
// thread's points so that independent points interleave in the FP32 pipes.  Parameters
// arrive as warp-uniform values decoded from the microcode.
#ifndef CC_OPS_CUH
#define CC_OPS_CUH

#include "cc_math.cuh"

// heavy, rarely dominant ops are real functions: inlining them PTS times per kernel variant
// would push the interpreter out of the instruction cache
#ifndef CC_OPT_NOINLINE_HEAVY
#define CC_OPT_NOINLINE_HEAVY 1
#endif
#if CC_OPT_NOINLINE_HEAVY
#define CC_DEV_HEAVY __device__ __noinline__
#else
#define CC_DEV_HEAVY __device__ __forceinline__
#endif

CC_DEV float4 cc_neg4(float4 a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
CC_DEV float4 cc_sel4(bool c, float4 a, float4 b)
{
    return make_float4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
}

// shapes/simple2d.cl:1-4 = perpendicular_intersection(slab_x, slab_y)  (common.cl:15-43)
CC_DEV float4 cc_rectangle(float hw, float hh, float4 p)
{
    float sx = copysignf(1.0f, p.x), sy = copysignf(1.0f, p.y);
    float ax = fabsf(p.x) - hw, ay = fabsf(p.y) - hh;
    float dist = cc_len2(ax, ay);
    float inv = cc_rcp(dist);
    bool both = ax > 0.0f && ay > 0.0f;
    bool first = ax > ay;
    float4 r;
    r.x = both ? sx * (ax * inv) : (first ? sx : 0.0f);
    r.y = both ? sy * (ay * inv) : (first ? 0.0f : sy);
    r.z = 0.0f;
    r.w = both ? dist : (first ? ax : ay);
    return r;
}

// shapes/simple3d.cl:18-21 = perpendicular_intersection(slab_z(h, coords), input)
CC_DEV float4 cc_extrusion(float h, float4 in, float cz)
{
    float sz = copysignf(1.0f, cz);
    float az = fabsf(cz) - h;
    float dist = cc_len2(az, in.w);
    float inv = cc_rcp(dist);
    float m1 = az * inv, m2 = in.w * inv;
    bool both = az > 0.0f && in.w > 0.0f;
    bool first = az > in.w;
    float4 r;
    r.x = both ? in.x * m2 : (first ? 0.0f : in.x);
    r.y = both ? in.y * m2 : (first ? 0.0f : in.y);
    r.z = both ? cc_fma(sz, m1, in.z * m2) : (first ? sz : in.z);
    r.w = both ? dist : (first ? az : in.w);
    return r;
}

// shapes/simple2d.cl:6-14
CC_DEV float4 cc_circle(float r, float4 p)
{
    float len = cc_len2(p.x, p.y);
    float inv = cc_rcp(len);
    bool zero = len == 0.0f;
    return make_float4(zero ? 1.0f : p.x * inv, zero ? 0.0f : p.y * inv, 0.0f, len - r);
}

// shapes/simple3d.cl:1-12
CC_DEV float4 cc_sphere(float r, float4 p)
{
    float len = cc_len3(p.x, p.y, p.z);
    float inv = cc_rcp(len);
    bool zero = len == 0.0f;
    return make_float4(zero ? 1.0f : p.x * inv, zero ? 0.0f : p.y * inv, zero ? 0.0f : p.z * inv, len - r);
}

// ---- the same ops over all PTS points of a thread, with one combined special-operand test ----
// Results are identical to the per-point forms above (which remain the slow path and the
// reference for the oracle): in the fast path every sqrt/rcp operand is in the exact range of
// cc_sqrt_fast / cc_rcp_fast, and a length in that range is never zero.
__device__ __noinline__ float4 cc_rectangle_slow(float hw, float hh, float4 p) { return cc_rectangle(hw, hh, p); }
__device__ __noinline__ float4 cc_circle_slow(float r, float4 p) { return cc_circle(r, p); }
__device__ __noinline__ float4 cc_sphere_slow(float r, float4 p) { return cc_sphere(r, p); }
__device__ __noinline__ float4 cc_extrusion_slow(float h, float4 in, float cz) { return cc_extrusion(h, in, cz); }

template <int PTS>
CC_DEV void cc_rectangle_n(float hw, float hh, float4 (&L)[PTS])
{
    float ax[PTS], ay[PTS], s[PTS];
    bool sp = false;
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        ax[j] = fabsf(L[j].x) - hw;
        ay[j] = fabsf(L[j].y) - hh;
        s[j] = cc_fma(ax[j], ax[j], ay[j] * ay[j]);
        sp |= cc_special(s[j]);
    }
    if (__any_sync(0xffffffffu, sp)) {  // warp-uniform: keeps the interpreter's control flow convergent
#pragma unroll
        for (int j = 0; j < PTS; ++j) L[j] = cc_rectangle_slow(hw, hh, L[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const float sx = copysignf(1.0f, L[j].x), sy = copysignf(1.0f, L[j].y);
        const float dist = cc_sqrt_fast(s[j]);
        const float inv = cc_rcp_fast(dist);
        const bool both = ax[j] > 0.0f && ay[j] > 0.0f;
        const bool first = ax[j] > ay[j];
        float4 r;
        r.x = both ? sx * (ax[j] * inv) : (first ? sx : 0.0f);
        r.y = both ? sy * (ay[j] * inv) : (first ? 0.0f : sy);
        r.z = 0.0f;
        r.w = both ? dist : (first ? ax[j] : ay[j]);
        L[j] = r;
    }
}

template <int PTS>
CC_DEV void cc_circle_n(float r, float4 (&L)[PTS])
{
    float s[PTS];
    bool sp = false;
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        s[j] = cc_fma(L[j].x, L[j].x, L[j].y * L[j].y);
        sp |= cc_special(s[j]);
    }
    if (__any_sync(0xffffffffu, sp)) {  // warp-uniform: keeps the interpreter's control flow convergent
#pragma unroll
        for (int j = 0; j < PTS; ++j) L[j] = cc_circle_slow(r, L[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const float len = cc_sqrt_fast(s[j]);
        const float inv = cc_rcp_fast(len);
        L[j] = make_float4(L[j].x * inv, L[j].y * inv, 0.0f, len - r);
    }
}

template <int PTS>
CC_DEV void cc_sphere_n(float r, float4 (&L)[PTS])
{
    float s[PTS];
    bool sp = false;
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        s[j] = cc_fma(L[j].x, L[j].x, cc_fma(L[j].y, L[j].y, L[j].z * L[j].z));
        sp |= cc_special(s[j]);
    }
    if (__any_sync(0xffffffffu, sp)) {  // warp-uniform: keeps the interpreter's control flow convergent
#pragma unroll
        for (int j = 0; j < PTS; ++j) L[j] = cc_sphere_slow(r, L[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const float len = cc_sqrt_fast(s[j]);
        const float inv = cc_rcp_fast(len);
        L[j] = make_float4(L[j].x * inv, L[j].y * inv, L[j].z * inv, len - r);
    }
}

// cz[j] = z coordinate of the point operand
template <int PTS>
CC_DEV void cc_extrusion_n(float h, float4 (&L)[PTS], const float (&cz)[PTS])
{
    float az[PTS], s[PTS];
    bool sp = false;
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        az[j] = fabsf(cz[j]) - h;
        s[j] = cc_fma(az[j], az[j], L[j].w * L[j].w);
        sp |= cc_special(s[j]);
    }
    if (__any_sync(0xffffffffu, sp)) {  // warp-uniform: keeps the interpreter's control flow convergent
#pragma unroll
        for (int j = 0; j < PTS; ++j) L[j] = cc_extrusion_slow(h, L[j], cz[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const float4 in = L[j];
        const float sz = copysignf(1.0f, cz[j]);
        const float dist = cc_sqrt_fast(s[j]);
        const float inv = cc_rcp_fast(dist);
        const float m1 = az[j] * inv, m2 = in.w * inv;
        const bool both = az[j] > 0.0f && in.w > 0.0f;
        const bool first = az[j] > in.w;
        float4 r;
        r.x = both ? in.x * m2 : (first ? 0.0f : in.x);
        r.y = both ? in.y * m2 : (first ? 0.0f : in.y);
        r.z = both ? cc_fma(sz, m1, in.z * m2) : (first ? sz : in.z);
        r.w = both ? dist : (first ? az[j] : in.w);
        L[j] = r;
    }
}

// shapes/common.cl:45-64
CC_DEV float4 cc_rounded_union(float r, float4 o1, float4 o2)
{
    float4 res = (o1.w < o2.w) ? o1 : o2;
    float c = cc_dot3(o1.x, o1.y, o1.z, o2.x, o2.y, o2.z);
    float x1 = r - o1.w, x2 = r - o2.w;
    if (c * x1 < x2 && c * x2 < x1) {  // rare: only within r of both surfaces
        float num = cc_fma(-((2.0f * c) * x1), x2, cc_fma(x1, x1, x2 * x2));
        float den = cc_fma(-c, c, 1.0f);
        res = make_float4(0.0f, 0.0f, 0.0f, r - cc_sqrt(cc_div(num, den)));
    }
    return res;
}

// 3x3 (row-major) * v + o   — common.cl:78-98 in matrix form
CC_DEV float4 cc_transform(const float (&m)[12], float x, float y, float z)
{
    return make_float4(cc_fma(m[0], x, cc_fma(m[1], y, cc_fma(m[2], z, m[9]))),
                       cc_fma(m[3], x, cc_fma(m[4], y, cc_fma(m[5], z, m[10]))),
                       cc_fma(m[6], x, cc_fma(m[7], y, cc_fma(m[8], z, m[11]))), 0.0f);
}

// common.cl:100-110 (matrix already divided by |q|^2; m[9] = |q|^2)
CC_DEV float4 cc_transform_from(const float (&m)[12], float4 in)
{
    return make_float4(cc_fma(m[0], in.x, cc_fma(m[1], in.y, m[2] * in.z)),
                       cc_fma(m[3], in.x, cc_fma(m[4], in.y, m[5] * in.z)),
                       cc_fma(m[6], in.x, cc_fma(m[7], in.y, m[8] * in.z)), in.w * m[9]);
}

// shapes/simple2d.cl:16-46; k = (piOverN, r, r*sin, r*cos, 2*piOverN)
CC_DEV_HEAVY float4 cc_regular_polygon2d(float k0, float k1, float k2, float k3, float k4, float4 co)
{
    const float k[5] = {k0, k1, k2, k3, k4};
    float piOverN = k[0], r = k[1];
    float len = cc_len2(co.x, co.y);
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)floorf(cc_div(alpha, k[4]));
    float t = (float)(side * 2) * piOverN;
    float modAlpha = (alpha - t) - piOverN;
    float s, c;
    cc_sincos(modAlpha, &s, &c);
    if (fabsf(s * len) > k[2]) {
        float ny, nx;
        cc_sincos(cc_fma(cc_sign(s), piOverN, t), &ny, &nx);
        float dx = co.x - nx * r, dy = co.y - ny * r;
        float dist = cc_len2(dx, dy);
        if (dist > 0.0f) {
            float inv = cc_rcp(dist);
            return make_float4(dx * inv, dy * inv, 0.0f, dist);
        }
    }
    float dy, dx;
    cc_sincos(t, &dy, &dx);
    return make_float4(dx, dy, 0.0f, cc_fma(len, c, -k[3]));
}

// shapes/gears.cl:1-42; k = (baseRadius, toothAngle, halfToothBaseAngle, 2*toothAngle, baseRadius^2)
CC_DEV_HEAVY float4 cc_involute_gear(float k0, float k1, float k2, float k3, float k4, float4 co)
{
    const float k[5] = {k0, k1, k2, k3, k4};
    float baseRadius = k[0], toothAngle = k[1], halfTooth = k[2];
    float len = cc_len2(co.x, co.y);
    float alpha = cc_atan2(co.y, co.x);
    float wrapped = cc_fmod_pos(alpha + CC_2PI_F, k[3]);
    float d = fabsf(wrapped - toothAngle);
    float involuteAlpha = halfTooth - d;
    if (len < baseRadius) {
        float inv = cc_rcp(len);
        float nx = co.y * inv, ny = -(co.x * inv);
        if (wrapped > toothAngle) { nx = -nx; ny = -ny; }
        return make_float4(nx, ny, 0.0f, (d - halfTooth) * len);
    }
    float phi = involuteAlpha + cc_acos(cc_div(baseRadius, len));
    float base = alpha - involuteAlpha;
    float normalAngle = (wrapped < toothAngle) ? (CC_PI_F - phi) - base : phi - base;
    float nx, ny;
    cc_sincos(normalAngle, &nx, &ny);
    float distance = cc_fma(-baseRadius, phi, cc_sqrt(cc_fma(len, len, -k[4])));
    return make_float4(nx, ny, 0.0f, distance);
}

// shapes/simple3d.cl:42-51
CC_DEV_HEAVY float4 cc_twist_revolution_to(float r, float twist, float4 co)
{
    float alpha = cc_fmod_pos(cc_atan2(co.z, co.x) + CC_PI_F, CC_2PI_F);
    float beta = cc_div(twist * alpha, CC_2PI_F);
    float ipx = cc_len2(co.x, co.z) - r, ipy = co.y;
    float s, c;
    cc_sincos(-beta, &s, &c);
    return make_float4(cc_fma(c, ipx, -(s * ipy)), cc_fma(s, ipx, c * ipy), 0.0f, 0.0f);
}

// shapes/simple3d.cl:53-97; k = (minorR, r, twist, min(1, lipschitz), padding)
CC_DEV_HEAVY float4 cc_twist_revolution_from(float k0, float k1, float k2, float k3, float k4, float4 inPlane,
                                             float4 co)
{
    const float k[5] = {k0, k1, k2, k3, k4};
    float minorR = k[0], r = k[1], twist = k[2];
    float ad = cc_len2(co.x, co.z);
    float ipx = ad - r, ipy = co.y;
    float icd = cc_len2(ipx, ipy);
    float wd = icd - minorR;
    float bound, dx, dy;
    if (ad == 0.0f) return make_float4(1.0f, 0.0f, 0.0f, r - minorR);
    if (wd > k[4]) {
        float inv = cc_rcp(icd);
        bound = wd; dx = ipx * inv; dy = ipy * inv;
    } else {
        float alpha = cc_fmod_pos(cc_atan2(co.z, co.x) + CC_PI_F, CC_2PI_F);
        float beta = cc_div(twist * alpha, CC_2PI_F);
        float s, c;
        cc_sincos(beta, &s, &c);
        bound = inPlane.w * k[3];
        dx = cc_fma(c, inPlane.x, -(s * inPlane.y));
        dy = cc_fma(s, inPlane.x, c * inPlane.y);
    }
    float mult = cc_div(dx, ad);
    return make_float4(co.x * mult, dy, co.z * mult, bound);
}

// shapes/unsafe.cl:8-15
CC_DEV_HEAVY float4 cc_circular_repetition_to(float piOverN, float twoPiOverN, float4 co)
{
    float len = cc_len2(co.x, co.y);
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)floorf(cc_div(alpha, twoPiOverN));
    float modAlpha = (alpha - (float)(side * 2) * piOverN) - piOverN;
    float s, c;
    cc_sincos(modAlpha, &s, &c);
    return make_float4(len * c, len * s, co.z, 0.0f);
}

// shapes/unsafe.cl:17-23
CC_DEV_HEAVY float4 cc_circular_repetition_from(float piOverN, float twoPiOverN, float4 dist, float4 co)
{
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)floorf(cc_div(alpha, twoPiOverN));
    float s, c;
    cc_sincos((float)(side * 2) * piOverN, &s, &c);
    return make_float4(cc_fma(c, dist.x, -(s * dist.y)), cc_fma(s, dist.x, c * dist.y), dist.z, dist.w);
}

// shapes/simple3d.cl:28-39
CC_DEV float4 cc_revolution_from(float4 flat, float4 co)
{
    float len = cc_len2(co.x, co.z);
    bool zero = len == 0.0f;
    float mult = zero ? flat.x : cc_div(flat.x, len);
    float cx = zero ? 1.0f : co.x;
    return make_float4(cx * mult, flat.y, co.z * mult, flat.w);
}

// Fused primitive (loader pattern initial_transformation_to -> [store p] -> circle|rectangle ->
// extrusion p -> [offset] -> [transformation_from]); bit-identical to the unfused sequence.
template <bool RECT, int PTS>
CC_DEV void cc_prim_n(const float (&m)[12], const float (&mf)[12], float pa, float pb, float h, float d,
                      const float (&x)[PTS], const float (&y)[PTS], const float (&z)[PTS], float4 (&L)[PTS])
{
    float pz[PTS];
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        L[j] = cc_transform(m, x[j], y[j], z[j]);
        pz[j] = L[j].z;
    }
    if (RECT) cc_rectangle_n<PTS>(pa, pb, L);
    else cc_circle_n<PTS>(pa, L);
    cc_extrusion_n<PTS>(h, L, pz);
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        L[j].w = L[j].w - d;
        L[j] = cc_transform_from(mf, L[j]);
    }
}

// shapes/polygons2d.cl:1-74 over a precomputed edge table (px, py, dx, dy, 1/|d|^2, cy) per
// edge; FETCH(i) returns table word i (the interpreter reads it from the microcode, the
// specialised kernels from a __constant__ array).
template <class FETCH>
CC_DEV float4 cc_polygon2d_core(const FETCH &fetch, uint32_t n, float4 co)
{
    float nnx = 0.0f, nny = 0.0f, nearest = __int_as_float(0x7f800000), outside = 1.0f;
    bool nearest_is_vertex = false;
    uint32_t e = 0;
    for (uint32_t i = 0; i < n; ++i, e += 6) {
        float px = fetch(e), py = fetch(e + 1), dx = fetch(e + 2), dy = fetch(e + 3), inv = fetch(e + 4),
              cy = fetch(e + 5);
        float tqx = co.x - px, tqy = co.y - py;
        float snx = -dy, sny = dx;
        if (((py < co.y) != (cy < co.y)) && (dy * cc_fma(snx, tqx, sny * tqy) > 0.0f)) outside = -outside;
        float t = cc_fma(dx, tqx, dy * tqy) * inv;
        if (t > 1.0f) continue;
        float cnx, cny, cd;
        bool civ;
        if (t >= 0.0f) {
            float tcx = cc_fma(-t, dx, tqx), tcy = cc_fma(-t, dy, tqy);
            cd = cc_fma(tcx, tcx, tcy * tcy);
            cnx = snx; cny = sny; civ = false;
        } else {
            cnx = tqx; cny = tqy;
            cd = cc_fma(cnx, cnx, cny * cny);
            civ = cd > 1.1920928955078125e-7f;
            if (!civ) { cnx = snx; cny = sny; }
        }
        if (cd < nearest) { nearest = cd; nnx = cnx; nny = cny; nearest_is_vertex = civ; }
    }
    float distance = outside * cc_sqrt(nearest);
    float inv = nearest_is_vertex ? cc_rcp(distance) : cc_rcp(cc_len2(nnx, nny));
    return make_float4(nnx * inv, nny * inv, 0.0f, distance);
}

struct cc_table_fetch {
    const float *t;
    CC_DEV float operator()(uint32_t i) const { return t[i]; }
};
__device__ __noinline__ float4 cc_polygon2d_table(const float *table, uint32_t n, float4 co)
{
    return cc_polygon2d_core(cc_table_fetch{table}, n, co);
}


#endif
