// TEST INFRASTRUCTURE — just enough of OpenCL C, as C++, to compile the reference's own
// device sources for the host (oracle/build_ref.py).  Not product code.
//
// build_ref.py takes the OpenCL program text the reference would hand to
// pyopencl.Program (cl_util/opencl_manager.py:116-141, gathered by importing the reference
// with a stub pyopencl), rewrites the three OpenCL-only syntaxes that C++ cannot parse
// (vector literals `(float4)(...)`, unsuffixed float constants under
// -cl-single-precision-constant, address-space qualifiers) and compiles it inside
// namespace clref against this header.  The built-in math maps to glibc's fp32 functions
// with no FMA contraction: one legitimate OpenCL implementation among many (the reference
// asks for -cl-fast-relaxed-math, so its last bits are implementation-defined anyway).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>

namespace clref {

typedef unsigned int uint;
typedef unsigned char uchar;

struct float2; struct float3; struct float4;

// ---- swizzle proxies (members of the unions below) ----
template <class V, int N, int A, int B, int C = 0>
struct swz {
    float d[4];
    operator V() const { V r; r.d[0] = d[A]; r.d[1] = d[B]; if (N > 2) r.d[2] = d[C]; return r; }
    swz &operator=(const V &v) { d[A] = v.d[0]; d[B] = v.d[1]; if (N > 2) d[C] = v.d[2]; return *this; }
};

struct float2 {
    union { float d[2]; struct { float x, y; }; struct { float s0, s1; }; };
    float2() : x(0), y(0) {}
    float2(float a) : x(a), y(a) {}  // OpenCL widens scalars implicitly
    float2(float a, float b) : x(a), y(b) {}
};
struct float3 {
    union {
        float d[4];
        struct { float x, y, z; };
        swz<float2, 2, 0, 1> xy;
        swz<float2, 2, 0, 2> xz;
    };
    float3() : x(0), y(0), z(0) {}
    explicit float3(float a) : x(a), y(a), z(a) {}  // (float3)scalar
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct float4 {
    union {
        float d[4];
        struct { float x, y, z, w; };
        swz<float2, 2, 0, 1> xy;
        swz<float2, 2, 0, 2> xz;
        swz<float3, 3, 0, 1, 2> xyz;
        swz<float3, 3, 1, 2, 3> yzw;
    };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float e) : x(a), y(b), z(c), w(e) {}
};
struct uint3 { uint x, y, z; };
struct uchar4 { uchar x, y, z, w; };
struct uint2 { uint x, y; };
struct int2;
struct int2_yx {  // the .yx swizzle of an int2 (rendering/polygon2d.cl)
    int d[2];
    operator int2() const;
};
struct int2 {
    union { int d[2]; struct { int x, y; }; int2_yx yx; };
    int2() : x(0), y(0) {}
    int2(int a, int b) : x(a), y(b) {}
};
inline int2_yx::operator int2() const { return int2(d[1], d[0]); }
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 &operator+=(int2 &a, int2 b) { a.x += b.x; a.y += b.y; return a; }
inline uint2 operator+(uint2 a, uint2 b) { uint2 r = {a.x + b.x, a.y + b.y}; return r; }

// ---- vector literals: (float4)(a, b, c, d) is rewritten to mk_float4(a, b, c, d) ----
inline float2 mk_float2(float a, float b) { return float2(a, b); }
inline float3 mk_float3(float a, float b, float c) { return float3(a, b, c); }
inline float4 mk_float4(float a, float b, float c, float d) { return float4(a, b, c, d); }
inline float4 mk_float4(float2 a, float c, float d) { return float4(a.x, a.y, c, d); }
inline float4 mk_float4(float3 a, float d) { return float4(a.x, a.y, a.z, d); }
inline float4 mk_float4(float a, float3 b) { return float4(a, b.x, b.y, b.z); }
inline uint3 mk_uint3(uint a, uint b, uint c) { uint3 r = {a, b, c}; return r; }
inline uint2 mk_uint2(uint a, uint b) { uint2 r = {a, b}; return r; }
inline int2 mk_int2(int a, int b) { return int2(a, b); }
inline float3 mk_float3(float a) { return float3(a, a, a); }
inline uchar4 mk_uchar4(uint a, uint b, uint c, uint d) { uchar4 r = {(uchar)a, (uchar)b, (uchar)c, (uchar)d}; return r; }

// ---- arithmetic ----
#define CLREF_OPS(V, N)                                                                            \
    inline V operator+(V a, V b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; } \
    inline V operator-(V a, V b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; } \
    inline V operator*(V a, V b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.d[i]; return r; } \
    inline V operator*(V a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }  \
    inline V operator*(float s, V a) { V r; for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }  \
    inline V operator/(V a, float s) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / s; return r; }  \
    inline V operator-(V a) { V r; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }              \
    inline V &operator*=(V &a, float s) { for (int i = 0; i < N; ++i) a.d[i] *= s; return a; }           \
    inline V &operator/=(V &a, float s) { for (int i = 0; i < N; ++i) a.d[i] /= s; return a; }           \
    inline V &operator-=(V &a, V b) { for (int i = 0; i < N; ++i) a.d[i] -= b.d[i]; return a; }          \
    inline V &operator+=(V &a, V b) { for (int i = 0; i < N; ++i) a.d[i] += b.d[i]; return a; }
CLREF_OPS(float2, 2)
CLREF_OPS(float3, 3)
CLREF_OPS(float4, 4)

inline float3 as_float3(float4 v) { return float3(v.x, v.y, v.z); }
inline float4 as_float4(float3 v) { return float4(v.x, v.y, v.z, 0.0f); }
inline float3 convert_float3(uint3 v) { return float3((float)v.x, (float)v.y, (float)v.z); }
inline float2 convert_float2(uint2 v) { return float2((float)v.x, (float)v.y); }
inline float3 operator+(float3 a, float s) { return float3(a.x + s, a.y + s, a.z + s); }
inline float3 operator*(int s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }

// ---- built-in math (fp32 glibc, no contraction) ----
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float3 cross(float3 a, float3 b)
{
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqrt(float x) { return ::sqrtf(x); }
inline float length(float2 v) { return ::sqrtf(dot(v, v)); }
inline float length(float3 v) { return ::sqrtf(dot(v, v)); }
inline float2 normalize(float2 v) { return v / length(v); }
inline float hypot(float a, float b) { return ::sqrtf(a * a + b * b); }
inline float fabs(float x) { return ::fabsf(x); }
inline float copysign(float a, float b) { return ::copysignf(a, b); }
inline float floor(float x) { return ::floorf(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float fmod(float a, float b) { return ::fmodf(a, b); }
inline float remainder(float a, float b) { return ::remainderf(a, b); }
inline float sincos(float x, float *c) { *c = ::cosf(x); return ::sinf(x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
// OpenCL leaves min/max of a NaN undefined; GPU implementations map them to the hardware's
// NaN-ignoring min/max instructions (PTX min.f32 / max.f32), which is fminf/fmaxf.  It matters in
// one place: the false-colour picture, where rays that miss everything carry a NaN normal.
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float2 vload2(size_t i, const float *p) { return float2(p[2 * i], p[2 * i + 1]); }
inline float3 normalize(float3 v) { return v / length(v); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a + (b - a) * t; }
inline float3 mix(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x)
{
    const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float3 fract(float3 v, float3 *ip)
{
    float3 r;
    for (int i = 0; i < 3; ++i) {
        ip->d[i] = ::floorf(v.d[i]);
        r.d[i] = ::fminf(v.d[i] - ip->d[i], 0x1.fffffep-1f);
    }
    return r;
}
inline float pow(float a, float b) { return ::powf(a, b); }
inline float exp(float a) { return ::expf(a); }
inline float fmin(float a, float b) { return ::fminf(a, b); }
inline float fmax(float a, float b) { return ::fmaxf(a, b); }

#define M_PI_F 3.14159274101257324f
#define M_PI_2_F 1.57079637050628662f
#undef M_PI
#define M_PI 3.14159274101257324f   /* -cl-single-precision-constant */
#define MAXFLOAT FLT_MAX

// ---- execution model: one work-item at a time, work-group size 1 ----
#define __kernel
#define __global
#define __constant const
#define __local
#define restrict __restrict__
#define CLK_LOCAL_MEM_FENCE 0
inline void barrier(int) {}
struct ndrange { size_t id[3]; size_t size[3]; };
extern thread_local ndrange g_nd;
inline size_t get_global_id(int i) { return g_nd.id[i]; }
inline size_t get_global_size(int i) { return g_nd.size[i]; }
inline size_t get_local_id(int) { return 0; }
inline uint atomic_inc(uint *p) { return __atomic_fetch_add(p, 1u, __ATOMIC_RELAXED); }
inline uint atomic_add(uint *p, uint v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

}  // namespace clref
