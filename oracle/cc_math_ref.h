/* TEST INFRASTRUCTURE — CPU oracle math layer.  Not product code.
 *
 * The reference builds its OpenCL with -cl-fast-relaxed-math
 * (/root/reference/codecad/cl_util/opencl_manager.py:12-18), so the vendor's
 * libm and FMA contraction decide the last bits of every result: bit-level
 * behaviour of the reference is implementation-defined.  DESIGN.md §"cc-arith"
 * therefore fixes ONE canonical fp32 arithmetic, and this header is its plain-C
 * restatement.  Rules:
 *   - every +,-,* is an IEEE-754 binary32 round-to-nearest-even operation,
 *     nothing is contracted implicitly (compile with -ffp-contract=off);
 *     fused multiply-adds are written explicitly as fmaf();
 *   - 1/x, x/y and sqrt are the correctly rounded IEEE operations;
 *   - transcendental functions are the fixed polynomial algorithms below
 *     (they use only the operations above, so any IEEE machine reproduces them).
 * Accuracy of the polynomials against glibc is checked by
 * tests/test_oracle_math.py (device-free).
 */
#ifndef CC_MATH_REF_H
#define CC_MATH_REF_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define CC_PI_F 3.14159274101257324f      /* (float)M_PI   */
#define CC_2PI_F 6.28318548202514648f     /* 2 * (float)M_PI, exact doubling */
#define CC_PI_2_F 1.57079637050628662f    /* (float)M_PI_2 */

static inline float cc_rcp(float x) { return 1.0f / x; }
static inline float cc_div(float x, float y) { return x / y; }
static inline float cc_sqrt(float x) { return sqrtf(x); }
static inline float cc_fma(float a, float b, float c) { return fmaf(a, b, c); }

static inline float cc_len2(float x, float y) { return cc_sqrt(cc_fma(x, x, y * y)); }
static inline float cc_len3(float x, float y, float z)
{
    return cc_sqrt(cc_fma(x, x, cc_fma(y, y, z * z)));
}
static inline float cc_dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return cc_fma(ax, bx, cc_fma(ay, by, az * bz));
}

/* floor / rint are exact operations in IEEE arithmetic. */
static inline float cc_floor(float x) { return floorf(x); }
static inline float cc_rint(float x) { return rintf(x); } /* ties-to-even (default mode) */

/* OpenCL sign(): +-1, or the (signed) zero itself, 0 for NaN. */
static inline float cc_sign(float x)
{
    if (x > 0.0f) return 1.0f;
    if (x < 0.0f) return -1.0f;
    if (x == 0.0f) return x;
    return 0.0f;
}

/* fmod for x >= 0, y > 0 (the only way the reference uses it:
 * simple3d.cl:44,83 and gears.cl:11).  q = floor(x/y), r = x - q*y with one
 * rounding, then one conditional correction each way so that 0 <= r < y even
 * when x/y rounded across an integer. */
static inline float cc_fmod_pos(float x, float y)
{
    float q = cc_floor(cc_div(x, y));
    float r = cc_fma(-q, y, x);
    if (r < 0.0f) r = r + y;
    if (r >= y) r = r - y;
    return r;
}

/* IEEE-style remainder: x - rint(x/y)*y.  remainder(x, +-inf) == x is relied on
 * by unsafe.cl:1-6 (Repetition with spacing None -> inf, shapes/unsafe.py:29-31). */
static inline float cc_remainder(float x, float y)
{
    if (isinf(y)) return x;
    float q = cc_rint(cc_div(x, y));
    return cc_fma(-q, y, x);
}

/* atan(a)/a on [0,1] as a degree-8 polynomial in a*a (max abs error 1.2e-8 in
 * exact arithmetic; coefficients from a Chebyshev-node fit, tools/fit_atan.py). */
static inline float cc_atan_unit(float a)
{
    float s = a * a;
    float p = 0.002834064298070311f;
    p = cc_fma(p, s, -0.016005030500026145f);
    p = cc_fma(p, s, 0.042587607460110644f);
    p = cc_fma(p, s, -0.07495445442927381f);
    p = cc_fma(p, s, 0.10636754097968429f);
    p = cc_fma(p, s, -0.14202570511671772f);
    p = cc_fma(p, s, 0.19992483578499645f);
    p = cc_fma(p, s, -0.33333066780691567f);
    p = cc_fma(p, s, 0.9999999842426363f);
    return p * a;
}

/* atan2(y, x).  atan2(0, 0) is defined as 0 (the reference only hits it on an
 * axis where the caller's result does not depend on the angle). */
static inline float cc_atan2(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float mx = ax > ay ? ax : ay;
    float mn = ax > ay ? ay : ax;
    float a = (mx == 0.0f) ? 0.0f : cc_div(mn, mx);
    float r = cc_atan_unit(a);
    if (ay > ax) r = CC_PI_2_F - r;
    if (x < 0.0f) r = CC_PI_F - r;
    return (y < 0.0f) ? -r : r;
}

/* sincos by Cody-Waite reduction to [-pi/4, pi/4] (3-term pi/2 split, good for
 * |x| < ~1e5, far beyond any angle the op library produces) and the Cephes
 * single-precision minimax polynomials. */
static inline void cc_sincos(float x, float *s_out, float *c_out)
{
    float k = cc_rint(x * 0.636619772367581343f); /* x * 2/pi */
    float r = cc_fma(-k, 1.5703125f, x);                         /* pi/2 hi  */
    r = cc_fma(-k, 4.83751296997070312e-4f, r);                  /* pi/2 mid */
    r = cc_fma(-k, 7.54978995489188194e-8f, r);                  /* pi/2 lo  */
    float z = r * r;
    float sp = -1.9515295891e-4f;
    sp = cc_fma(sp, z, 8.3321608736e-3f);
    sp = cc_fma(sp, z, -1.6666654611e-1f);
    float s = cc_fma(sp * z, r, r);
    float cp = 2.443315711809948e-5f;
    cp = cc_fma(cp, z, -1.388731625493765e-3f);
    cp = cc_fma(cp, z, 4.166664568298827e-2f);
    float c = cc_fma(cp * z, z, cc_fma(-0.5f, z, 1.0f));
    int q = (int)k & 3;
    float ss = (q & 1) ? c : s;
    float cc = (q & 1) ? s : c;
    if (q & 2) ss = -ss;
    if ((q + 1) & 2) cc = -cc;
    *s_out = ss;
    *c_out = cc;
}

/* acos(x) for x in [-1, 1] through atan2(sqrt(1 - x*x), x). */
static inline float cc_acos(float x)
{
    float t = cc_fma(-x, x, 1.0f);
    if (t < 0.0f) t = 0.0f;
    return cc_atan2(cc_sqrt(t), x);
}

#endif /* CC_MATH_REF_H */
