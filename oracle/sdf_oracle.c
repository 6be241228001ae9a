/* TEST INFRASTRUCTURE — CPU oracle for the codecad SDF hot path.  Not product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * It is a plain-C restatement of the reference's OpenCL device code, consuming the
 * reference's own float32 program words:
 *   interpreter  /root/reference/codecad/nodes/codegen.py:17-63 (+ handlers :91-134)
 *   opcodes      /root/reference/codecad/nodes/node.py:12-56
 *   encoding     /root/reference/codecad/nodes/program.py:39-71
 *   op library   /root/reference/codecad/shapes/{common,simple2d,simple3d,polygons2d,unsafe,gears}.cl
 *                /root/reference/codecad/cl_util/util.cl:1-15
 *   kernels      /root/reference/codecad/grid_eval.cl:2-34, subdivision.cl:12-30,
 *                mass_properties.cl:7-56, cl_util/indexing.h:4
 *
 * Arithmetic: the canonical "cc-arith" of DESIGN.md (see cc_math_ref.h).  Parameter-
 * only sub-expressions (quaternion -> 3x3 matrix, polygon edge tables, gear
 * constants) are evaluated once per instruction, in double where stated, exactly as
 * the device program loader does; everything per point is fp32.
 *
 * Pinning: no OpenCL runtime exists in the build container or on the GPU box, so
 * bit-level parity with a vendor OpenCL build is UNPINNED.  The oracle is pinned
 * (0) against the reference's 32 golden pictures (tests/baseline/rendered_*.png, copied to
 * tests/golden/reference_renders): the renderers restated at the end of this file reproduce every one
 * within the reference's own tolerance (tests/test_oracle_golden_images.py),
 * (a) against every known-answer test the reference holds for this path
 * (tests/test_oracle_known_answers.py: tests/test_mass_properties.py:16-108,
 * tests/test_subdivision.py:110-161, tests/test_dsdf.py:113-192 of the reference) and
 * (b) against oracle/_ref, the reference's own .cl sources compiled for the CPU
 * (oracle/build_ref.py), within the north-star tolerance.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cc_math_ref.h"

#define REGISTER_COUNT 512

typedef struct { float x, y, z, w; } v4;

enum {
    OP_RETURN = 0, OP_STORE, OP_LOAD, OP_RECTANGLE, OP_CIRCLE, OP_REGULAR_POLYGON2D,
    OP_POLYGON2D, OP_SPHERE, OP_HALF_SPACE, OP_REVOLUTION_TO, OP_TWIST_REVOLUTION_TO,
    OP_INITIAL_TRANSFORMATION_TO, OP_TRANSFORMATION_TO, OP_TRANSFORMATION_FROM,
    OP_MIRROR, OP_SYMMETRICAL_TO, OP_OFFSET, OP_SHELL, OP_REPETITION,
    OP_CIRCULAR_REPETITION_TO, OP_CIRCULAR_REPETITION_FROM, OP_INVOLUTE_GEAR,
    OP_EXTRUSION, OP_REVOLUTION_FROM, OP_TWIST_REVOLUTION_FROM, OP_SYMMETRICAL_FROM,
    OP_UNION, OP_INTERSECTION, OP_SUBTRACTION, OP_COUNT
};

/* parameter words per opcode (node.py:18-52); -1 = polygon2d (1 + 2n) */
static const int NPARAMS[OP_COUNT] = {0, 0, 0, 2, 1, 2, -1, 1, 0, 0, 2, 7, 7, 4, 0, 0,
                                      1, 1, 3, 1, 1, 2, 1, 0, 3, 0, 1, 1, 1};

typedef struct {
    int op;
    int reg;
    const float *p; /* parameters inside the caller's word array */
    float k[16];    /* per-instruction constants derived from the parameters */
    float *edges;   /* polygon2d: 5 floats per edge (px, py, dx, dy, 1/|d|^2) */
    float *groups;  /* polygon2d, product formulation only: per 8 edges (xmin, xmax, ymin, ymax) + 2 spare */
    int n;
} ins_t;

typedef struct {
    ins_t *ins;
    int count;
    long flops; /* not used by parity, kept for the bench's executed-flop estimate */
} prog_t;

/* ---- per-instruction constants (mirrors the device loader) -------------------- */

/* quaternion (x,y,z,w) -> row-major 3x3 of  v -> 2 v(v.p) + 2 w (v x p) + (w^2-v.v) p
 * (common.cl:1-6), in double, optionally divided by |q|^2 (common.cl:100-110). */
static void quat_matrix(const float *q, int divide_by_scale, float *m, float *scale)
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double k = w * w - (x * x + y * y + z * z);
    double d[9];
    d[0] = 2.0 * (x * x) + k;
    d[1] = 2.0 * (x * y - w * z);
    d[2] = 2.0 * (x * z + w * y);
    d[3] = 2.0 * (x * y + w * z);
    d[4] = 2.0 * (y * y) + k;
    d[5] = 2.0 * (y * z - w * x);
    d[6] = 2.0 * (x * z - w * y);
    d[7] = 2.0 * (y * z + w * x);
    d[8] = 2.0 * (z * z) + k;
    double s = (x * x + y * y) + (z * z + w * w);
    for (int i = 0; i < 9; ++i)
        m[i] = (float)(divide_by_scale ? d[i] / s : d[i]);
    if (scale) *scale = (float)s;
}

static int prepare(const float *words, int n_words, prog_t *prog)
{
    int cap = 64, count = 0, pc = 0;
    ins_t *ins = (ins_t *)calloc((size_t)cap, sizeof(ins_t));
    for (;;) {
        if (pc >= n_words) { free(ins); return -1; }
        float wf = words[pc];
        if (!(wf >= 0.0f) || wf >= (float)(OP_COUNT * REGISTER_COUNT)) { free(ins); return -2; }
        unsigned instruction = (unsigned)wf;
        int op = (int)(instruction / REGISTER_COUNT);
        int reg = (int)(instruction % REGISTER_COUNT);
        int np = NPARAMS[op];
        if (np < 0) {
            if (pc + 1 >= n_words) { free(ins); return -1; }
            np = 1 + 2 * (int)words[pc + 1];
        }
        if (pc + 1 + np > n_words) { free(ins); return -1; }
        if (count == cap) {
            cap *= 2;
            ins = (ins_t *)realloc(ins, (size_t)cap * sizeof(ins_t));
        }
        ins_t *I = &ins[count++];
        memset(I, 0, sizeof(*I));
        I->op = op;
        I->reg = reg;
        I->p = words + pc + 1;
        const float *p = I->p;
        switch (op) {
        case OP_INITIAL_TRANSFORMATION_TO:
        case OP_TRANSFORMATION_TO:
            quat_matrix(p, 0, I->k, NULL);
            I->k[9] = p[4]; I->k[10] = p[5]; I->k[11] = p[6];
            break;
        case OP_TRANSFORMATION_FROM:
            quat_matrix(p, 1, I->k, &I->k[9]);
            break;
        case OP_REGULAR_POLYGON2D:
            I->k[0] = p[1] * (float)sin((double)p[0]); /* r * sin(pi/n)  simple2d.cl:25 */
            I->k[1] = p[1] * (float)cos((double)p[0]); /* r * cos(pi/n)  simple2d.cl:45 */
            I->k[2] = 2.0f * p[0];
            break;
        case OP_CIRCULAR_REPETITION_TO:
        case OP_CIRCULAR_REPETITION_FROM:
            I->k[2] = 2.0f * p[0];
            break;
        case OP_INVOLUTE_GEAR: {
            float pa = p[1];
            I->k[0] = (float)cos((double)pa);                 /* baseRadius   gears.cl:2 */
            I->k[1] = cc_div(CC_PI_F, p[0]);                  /* toothAngle   gears.cl:3 */
            I->k[2] = (I->k[1] * 0.5f + (float)tan((double)pa)) - pa; /* gears.cl:6 */
            I->k[3] = 2.0f * I->k[1];
            I->k[4] = I->k[0] * I->k[0];
            break;
        }
        case OP_TWIST_REVOLUTION_FROM: {
            float minorR = p[0], r = p[1], twist = p[2];
            float arg = fminf(CC_PI_F, cc_div(CC_PI_2_F * CC_PI_2_F, fabsf(twist)));
            float lip = cc_div(((r - minorR) * 2.0f) * (float)sin((double)arg), minorR);
            I->k[0] = fminf(1.0f, lip);  /* simple3d.cl:86-88 */
            I->k[1] = 0.05f * r;         /* wrapperPadding simple3d.cl:70 */
            break;
        }
        case OP_POLYGON2D: {
            int n = (int)p[0];
            I->n = n;
            I->edges = (float *)malloc(sizeof(float) * 5 * (size_t)(n > 0 ? n : 1));
            for (int i = 0; i < n; ++i) {
                int j = (i + n - 1) % n; /* previous point: polygons2d.cl:13,17-18 */
                float px = p[1 + 2 * j], py = p[2 + 2 * j];
                float cx = p[1 + 2 * i], cy = p[2 + 2 * i];
                float dx = cx - px, dy = cy - py;
                float *e = I->edges + 5 * i;
                e[0] = px; e[1] = py; e[2] = dx; e[3] = dy;
                e[4] = cc_rcp(cc_fma(dx, dx, dy * dy));
            }
            I->groups = (float *)malloc(sizeof(float) * 6 * (size_t)((n + 7) / 8 + 1));
            for (int g0 = 0, g = 0; g0 < n; g0 += 8, ++g) {
                int j = (g0 + n - 1) % n;
                float xmin = p[1 + 2 * j], xmax = xmin, ymin = p[2 + 2 * j], ymax = ymin;
                for (int k = g0; k < n && k < g0 + 8; ++k) {
                    xmin = fminf(xmin, p[1 + 2 * k]); xmax = fmaxf(xmax, p[1 + 2 * k]);
                    ymin = fminf(ymin, p[2 + 2 * k]); ymax = fmaxf(ymax, p[2 + 2 * k]);
                }
                float *q = I->groups + 6 * g;
                q[0] = xmin; q[1] = xmax; q[2] = ymin; q[3] = ymax; q[4] = p[1 + 2 * j]; q[5] = p[2 + 2 * j];
            }
            break;
        }
        default:
            break;
        }
        pc += 1 + np;
        if (op == OP_RETURN) break;
    }
    prog->ins = ins;
    prog->count = count;
    return 0;
}

static void release(prog_t *prog)
{
    for (int i = 0; i < prog->count; ++i) { free(prog->ins[i].edges); free(prog->ins[i].groups); }
    free(prog->ins);
    prog->ins = NULL;
}

/* ---- executed-branch algorithmic flop counter ---------------------------------
 * Counting rules of SURVEY.md 8(a3): add/sub/mul/div/sqrt = 1, FMA = 2, compare/select/abs/neg free,
 * libm-class calls not counted; costs are those of the REFERENCE's formulation of each op (the
 * static minimum per op is the table the device loader uses for cc_program_info.flops_min).  The
 * counter adds, per evaluation, the cost of the branches that were actually taken. */
static __thread unsigned long long t_flops;
#define FL(n) (t_flops += (unsigned long long)(n))
static const unsigned char BASE_FLOPS[OP_COUNT] = {
    /* return store load */ 0, 0, 0, /* rectangle */ 2, /* circle */ 7, /* regular_polygon2d */ 25, /* polygon2d */ 10,
    /* sphere */ 11, /* half_space */ 0, /* revolution_to */ 4, /* twist_revolution_to */ 16,
    /* initial_transformation_to */ 42, /* transformation_to */ 42, /* transformation_from */ 50, /* mirror */ 0,
    /* symmetrical_to */ 0, /* offset */ 1, /* shell */ 1, /* repetition */ 3, /* circular_repetition_to */ 14,
    /* circular_repetition_from */ 14, /* involute_gear */ 20, /* extrusion */ 1, /* revolution_from */ 7,
    /* twist_revolution_from */ 15, /* symmetrical_from */ 0, /* union isect sub: 9 when rounded */ 0, 0, 0};

/* ---- op library ----------------------------------------------------------------- */

static inline v4 mk(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v4 neg(v4 a) { return mk(-a.x, -a.y, -a.z, -a.w); }

/* common.cl:15-31 */
static inline v4 perpendicular_intersection(v4 a, v4 b)
{
    if (a.w > 0.0f && b.w > 0.0f) {
        FL(15);
        float dist = cc_len2(a.w, b.w);
        float inv = cc_rcp(dist);
        float m1 = a.w * inv, m2 = b.w * inv;
        return mk(cc_fma(a.x, m1, b.x * m2), cc_fma(a.y, m1, b.y * m2),
                  cc_fma(a.z, m1, b.z * m2), dist);
    }
    return (a.w > b.w) ? a : b;
}

/* common.cl:33-43 */
static inline v4 slab_x(float h, v4 p) { return mk(copysignf(1.0f, p.x), 0, 0, fabsf(p.x) - h); }
static inline v4 slab_y(float h, v4 p) { return mk(0, copysignf(1.0f, p.y), 0, fabsf(p.y) - h); }
static inline v4 slab_z(float h, v4 p) { return mk(0, 0, copysignf(1.0f, p.z), fabsf(p.z) - h); }

/* common.cl:45-64.  cc-arith pins down two cases the formula leaves to rounding noise,
 * both no-ops in exact arithmetic where |c| <= 1: the blend needs at least one operand closer than r
 * (with both farther it would take c > 1), and the radicand is clamped at zero (never NaN). */
static inline v4 rounded_union(float r, v4 o1, v4 o2)
{
    if (r >= 0.0f) {
        float c = cc_dot3(o1.x, o1.y, o1.z, o2.x, o2.y, o2.z);
        float x1 = r - o1.w, x2 = r - o2.w;
        FL(9);
        if (c * x1 < x2 && c * x2 < x1 && (x1 > 0.0f || x2 > 0.0f)) {
            FL(12);
            float num = cc_fma(-((2.0f * c) * x1), x2, cc_fma(x1, x1, x2 * x2));
            float den = cc_fma(-c, c, 1.0f);
            float d = r - cc_sqrt(fmaxf(cc_div(num, den), 0.0f));
            return mk(0, 0, 0, d);
        }
    }
    return (o1.w < o2.w) ? o1 : o2;
}

/* cc-arith: a matrix row is accumulated innermost-first (z, y, x) and terms whose
 * coefficient is exactly zero are OMITTED — a parameter-only simplification fixed at load time,
 * like the quaternion -> matrix conversion itself.  (Multiplying by a zero coefficient would only
 * decide the sign of an exactly-zero result.) */
static inline float row_to(const float *m, float x, float y, float z, float o)
{
    float acc = o;
    if (m[2] != 0.0f) acc = cc_fma(m[2], z, acc);
    if (m[1] != 0.0f) acc = cc_fma(m[1], y, acc);
    if (m[0] != 0.0f) acc = cc_fma(m[0], x, acc);
    return acc;
}
static inline float row_from(const float *m, float x, float y, float z)
{
    float acc = 0.0f;
    int have = 0;
    if (m[2] != 0.0f) { acc = m[2] * z; have = 1; }
    if (m[1] != 0.0f) { acc = have ? cc_fma(m[1], y, acc) : m[1] * y; have = 1; }
    if (m[0] != 0.0f) { acc = have ? cc_fma(m[0], x, acc) : m[0] * x; have = 1; }
    return acc;
}
static inline v4 apply_matrix(const float *m, float x, float y, float z, float ox, float oy, float oz)
{
    return mk(row_to(m, x, y, z, ox), row_to(m + 3, x, y, z, oy), row_to(m + 6, x, y, z, oz), 0.0f);
}

/* simple2d.cl:16-46 */
static inline v4 regular_polygon2d(const ins_t *I, v4 co)
{
    float piOverN = I->p[0], r = I->p[1];
    float len = cc_len2(co.x, co.y);
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)cc_floor(cc_div(alpha, I->k[2]));
    float t = (float)(side * 2) * piOverN;
    float modAlpha = (alpha - t) - piOverN;
    float s, c;
    cc_sincos(modAlpha, &s, &c);
    if (fabsf(s * len) > I->k[0]) {
        FL(15);
        float ny, nx;
        cc_sincos(cc_fma(cc_sign(s), piOverN, t), &ny, &nx);
        float dx = co.x - nx * r, dy = co.y - ny * r;
        float dist = cc_len2(dx, dy);
        if (dist > 0.0f) {
            float inv = cc_rcp(dist);
            return mk(dx * inv, dy * inv, 0, dist);
        }
    }
    float dy, dx;
    cc_sincos(t, &dy, &dx);
    return mk(dx, dy, 0, cc_fma(len, c, -I->k[1]));
}

/* polygons2d.cl:1-74 */
static inline v4 polygon2d(const ins_t *I, v4 co)
{
    float nnx = 0.0f, nny = 0.0f, nearest = INFINITY, outside = 1.0f;
    int nearest_is_vertex = 0;
    for (int i = 0; i < I->n; ++i) {
        const float *e = I->edges + 5 * i;
        float px = e[0], py = e[1], dx = e[2], dy = e[3];
        float cy = I->p[2 + 2 * i];
        float tqx = co.x - px, tqy = co.y - py;
        float snx = -dy, sny = dx;
        if (((py < co.y) != (cy < co.y)) && (dy * cc_fma(snx, tqx, sny * tqy) > 0.0f))
            outside = -outside;
        float t = cc_fma(dx, tqx, dy * tqy) * e[4];
        FL(15 + (t > 1.0f ? 0 : (t >= 0.0f ? 7 : 5)));
        if (t > 1.0f) continue;
        float cnx, cny, cd;
        int civ;
        if (t >= 0.0f) {
            float tcx = cc_fma(-t, dx, tqx), tcy = cc_fma(-t, dy, tqy);
            cd = cc_fma(tcx, tcx, tcy * tcy);
            cnx = snx; cny = sny; civ = 0;
        } else {
            cnx = tqx; cny = tqy;
            cd = cc_fma(cnx, cnx, cny * cny);
            civ = cd > 1.1920928955078125e-7f; /* FLT_EPSILON */
            if (!civ) { cnx = snx; cny = sny; }
        }
        if (cd < nearest) { nearest = cd; nnx = cnx; nny = cny; nearest_is_vertex = civ; }
    }
    float distance = outside * cc_sqrt(nearest);
    float inv = nearest_is_vertex ? cc_rcp(distance) : cc_rcp(cc_len2(nnx, nny));
    return mk(nnx * inv, nny * inv, 0, distance);
}

/* The PRODUCT's formulation of the same op (codecad_b200/csrc/cc_ops.cuh cc_polygon2d_v), restated
 * here only so that tests/test_oracle_known_answers.py can compare the two formulations on millions
 * of points on the CPU: branch-free edge loop keeping the smallest squared distance and the index of
 * the edge that produced it, t clamped at zero instead of a separate start-vertex candidate, crossing
 * test reusing the previous end-point comparison; feature kind and normal recomputed after the loop. */
static int g_polygon_alt = 0;
static unsigned long long g_polygon_groups[3]; /* product formulation: edge groups seen / run for crossing only / run for distance */
void oracle_polygon_group_counters(unsigned long long out[3], int reset)
{
    for (int i = 0; i < 3; ++i) { out[i] = g_polygon_groups[i]; if (reset) g_polygon_groups[i] = 0; }
}
static inline v4 polygon2d_alt(const ins_t *I, v4 co)
{
    float nearest = INFINITY, best = -1.0f, outside = 1.0f;
    int prev_below = I->n ? (I->edges[1] < co.y) : 0;
    /* edge groups (cc_ops.cuh): B bounds the final nearest value from above; a group whose bounding box
     * is farther than B, and whose y range does not contain the point, cannot change a bit */
    const int ng = (I->n + 7) / 8;
    float bound = INFINITY;
    for (int i = 0; i < I->n; i += 4) { /* CC_POLY_BOUND_STRIDE */
        float qx = co.x - I->edges[5 * i], qy = co.y - I->edges[5 * i + 1];
        bound = fminf(bound, cc_fma(qx, qx, qy * qy));
    }
    bound = bound * 1.00002f;
    for (int g = 0; g < ng; ++g) {
        const float *q = I->groups + 6 * g;
        float ex = fmaxf(fmaxf(q[0] - co.x, co.x - q[1]), 0.0f), ey = fmaxf(fmaxf(q[2] - co.y, co.y - q[3]), 0.0f);
        float lb = cc_fma(ex, ex, ey * ey) * 0.99998f;
        int above = co.y > q[3];
        int may_cross = (co.y > q[2]) && !above;
        int may_win = !(lb > bound);
        __atomic_fetch_add(&g_polygon_groups[0], 1ull, __ATOMIC_RELAXED);
        if (!may_cross && !may_win) { prev_below = above; continue; }
        __atomic_fetch_add(&g_polygon_groups[1 + (may_win ? 1 : 0)], 1ull, __ATOMIC_RELAXED);
        if (!may_win) { /* crossing parity only */
            for (int i = 8 * g; i < I->n && i < 8 * g + 8; ++i) {
                const float *e = I->edges + 5 * i;
                float cy = I->p[2 + 2 * i];
                int cur_below = cy < co.y;
                if (prev_below != cur_below) {
                    float tqx = co.x - e[0], tqy = co.y - e[1];
                    float side = e[3] * cc_fma(-e[3], tqx, e[2] * tqy);
                    if (side > 0.0f) outside = -outside;
                }
                prev_below = cur_below;
            }
            continue;
        }
    for (int i = 8 * g; i < I->n && i < 8 * g + 8; ++i) {
        const float *e = I->edges + 5 * i;
        float px = e[0], py = e[1], dx = e[2], dy = e[3], cy = I->p[2 + 2 * i];
        float tqx = co.x - px, tqy = co.y - py;
        int cur_below = cy < co.y;
        float side = dy * cc_fma(-dy, tqx, dx * tqy);
        if ((prev_below != cur_below) && side > 0.0f) outside = -outside;
        prev_below = cur_below;
        float t = cc_fma(dx, tqx, dy * tqy) * e[4];
        FL(15 + (t > 1.0f ? 0 : (t >= 0.0f ? 7 : 5)));
        float tc = fmaxf(t, 0.0f);
        float tcx = cc_fma(-tc, dx, tqx), tcy = cc_fma(-tc, dy, tqy);
        float cd = cc_fma(tcx, tcx, tcy * tcy);
        int better = !(t > 1.0f) && (cd < nearest);
        nearest = better ? cd : nearest;
        best = better ? (float)i : best;
        (void)py;
    }
    }
    float nnx = 0.0f, nny = 0.0f;
    int nearest_is_vertex = 0;
    if (best >= 0.0f) {
        const float *e = I->edges + 5 * (int)best;
        float px = e[0], py = e[1], dx = e[2], dy = e[3];
        float tqx = co.x - px, tqy = co.y - py;
        float t = cc_fma(dx, tqx, dy * tqy) * e[4];
        nnx = -dy; nny = dx;
        if (!(t >= 0.0f)) {
            float cd = cc_fma(tqx, tqx, tqy * tqy);
            nearest_is_vertex = cd > 1.1920928955078125e-7f;
            if (nearest_is_vertex) { nnx = tqx; nny = tqy; }
        }
    }
    float distance = outside * cc_sqrt(nearest);
    float inv = nearest_is_vertex ? cc_rcp(distance) : cc_rcp(cc_len2(nnx, nny));
    return mk(nnx * inv, nny * inv, 0, distance);
}
void oracle_set_polygon_formulation(int alt) { g_polygon_alt = alt; }

/* gears.cl:1-42 */
static inline v4 involute_gear(const ins_t *I, v4 co)
{
    float baseRadius = I->k[0], toothAngle = I->k[1], halfTooth = I->k[2];
    float len = cc_len2(co.x, co.y);
    float alpha = cc_atan2(co.y, co.x);
    float wrapped = cc_fmod_pos(alpha + CC_2PI_F, I->k[3]);
    float d = fabsf(wrapped - toothAngle);
    float involuteAlpha = halfTooth - d;
    if (len < baseRadius) {
        float inv = cc_rcp(len);
        float nx = co.y * inv, ny = -(co.x * inv);
        if (wrapped > toothAngle) { nx = -nx; ny = -ny; }
        return mk(nx, ny, 0, (d - halfTooth) * len);
    }
    FL(10);
    float phi = involuteAlpha + cc_acos(cc_div(baseRadius, len));
    float base = alpha - involuteAlpha;
    float normalAngle = (wrapped < toothAngle) ? (CC_PI_F - phi) - base : phi - base;
    float nx, ny;
    cc_sincos(normalAngle, &nx, &ny);
    float distance = cc_fma(-baseRadius, phi, cc_sqrt(cc_fma(len, len, -I->k[4])));
    return mk(nx, ny, 0, distance);
}

/* simple3d.cl:42-51 */
static inline v4 twist_revolution_to(const ins_t *I, v4 co)
{
    float r = I->p[0], twist = I->p[1];
    float alpha = cc_fmod_pos(cc_atan2(co.z, co.x) + CC_PI_F, CC_2PI_F);
    float beta = cc_div(twist * alpha, CC_2PI_F);
    float ipx = cc_len2(co.x, co.z) - r, ipy = co.y;
    float s, c;
    cc_sincos(-beta, &s, &c);
    return mk(cc_fma(c, ipx, -(s * ipy)), cc_fma(s, ipx, c * ipy), 0, 0);
}

/* simple3d.cl:53-97 */
static inline v4 twist_revolution_from(const ins_t *I, v4 inPlane, v4 co)
{
    float minorR = I->p[0], r = I->p[1], twist = I->p[2];
    float ad = cc_len2(co.x, co.z);
    float ipx = ad - r, ipy = co.y;
    float icd = cc_len2(ipx, ipy);
    float wd = icd - minorR;
    float bound, dx, dy;
    if (ad == 0.0f) return mk(1, 0, 0, r - minorR);
    if (!(wd > I->k[1])) FL(25);
    if (wd > I->k[1]) {
        float inv = cc_rcp(icd);
        bound = wd; dx = ipx * inv; dy = ipy * inv;
    } else {
        float alpha = cc_fmod_pos(cc_atan2(co.z, co.x) + CC_PI_F, CC_2PI_F);
        float beta = cc_div(twist * alpha, CC_2PI_F);
        float s, c;
        cc_sincos(beta, &s, &c);
        bound = inPlane.w * I->k[0];
        dx = cc_fma(c, inPlane.x, -(s * inPlane.y));
        dy = cc_fma(s, inPlane.x, c * inPlane.y);
    }
    float mult = cc_div(dx, ad);
    return mk(co.x * mult, dy, co.z * mult, bound);
}

/* unsafe.cl:8-23 */
static inline v4 circular_repetition_to(const ins_t *I, v4 co)
{
    float piOverN = I->p[0];
    float len = cc_len2(co.x, co.y);
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)cc_floor(cc_div(alpha, I->k[2]));
    float modAlpha = (alpha - (float)(side * 2) * piOverN) - piOverN;
    float s, c;
    cc_sincos(modAlpha, &s, &c);
    return mk(len * c, len * s, co.z, 0);
}

static inline v4 circular_repetition_from(const ins_t *I, v4 dist, v4 co)
{
    float piOverN = I->p[0];
    float alpha = (cc_atan2(co.y, co.x) + CC_2PI_F) + piOverN;
    int side = (int)cc_floor(cc_div(alpha, I->k[2]));
    float s, c;
    cc_sincos((float)(side * 2) * piOverN, &s, &c);
    return mk(cc_fma(c, dist.x, -(s * dist.y)), cc_fma(s, dist.x, c * dist.y), dist.z, dist.w);
}

/* ---- interpreter: codegen.py:17-63 ---------------------------------------------- */

static v4 evaluate(const prog_t *prog, float px, float py, float pz)
{
    v4 registers[REGISTER_COUNT];
    v4 last = mk(0, 0, 0, 0);
    const ins_t *I = prog->ins;
    for (;; ++I) {
        const float *p = I->p;
        FL(BASE_FLOPS[I->op]);
#define in2 (registers[I->reg]) /* second operand of arity-2 ops / source of _load */
        switch (I->op) {
        case OP_RETURN: return last;
        case OP_STORE: registers[I->reg] = last; break;
        case OP_LOAD: last = in2; break;
        case OP_RECTANGLE: /* simple2d.cl:1-4 */
            last = perpendicular_intersection(slab_x(p[0], last), slab_y(p[1], last));
            break;
        case OP_CIRCLE: { /* simple2d.cl:6-14 */
            float len = cc_len2(last.x, last.y);
            if (len == 0.0f) last = mk(1, 0, 0, len - p[0]);
            else { float inv = cc_rcp(len); last = mk(last.x * inv, last.y * inv, 0, len - p[0]); }
            break;
        }
        case OP_REGULAR_POLYGON2D: last = regular_polygon2d(I, last); break;
        case OP_POLYGON2D: last = g_polygon_alt ? polygon2d_alt(I, last) : polygon2d(I, last); break;
        case OP_SPHERE: { /* simple3d.cl:1-12 */
            float len = cc_len3(last.x, last.y, last.z);
            if (len == 0.0f) last = mk(1, 0, 0, len - p[0]);
            else { float inv = cc_rcp(len); last = mk(last.x * inv, last.y * inv, last.z * inv, len - p[0]); }
            break;
        }
        case OP_HALF_SPACE: last = mk(0, -1, 0, -last.y); break; /* simple3d.cl:14-16 */
        case OP_REVOLUTION_TO: last = mk(cc_len2(last.x, last.z), last.y, 0, 0); break; /* :23-26 */
        case OP_TWIST_REVOLUTION_TO: last = twist_revolution_to(I, last); break;
        case OP_INITIAL_TRANSFORMATION_TO: /* common.cl:78-87 */
            last = apply_matrix(I->k, px, py, pz, I->k[9], I->k[10], I->k[11]);
            break;
        case OP_TRANSFORMATION_TO: /* common.cl:89-98 */
            last = apply_matrix(I->k, last.x, last.y, last.z, I->k[9], I->k[10], I->k[11]);
            break;
        case OP_TRANSFORMATION_FROM: { /* common.cl:100-110 */
            const float *m = I->k;
            float x = last.x, y = last.y, z = last.z;
            last = mk(row_from(m, x, y, z), row_from(m + 3, x, y, z), row_from(m + 6, x, y, z), last.w * I->k[9]);
            break;
        }
        case OP_MIRROR: last.x = -last.x; break;          /* common.cl:112-114 */
        case OP_SYMMETRICAL_TO: last.x = fabsf(last.x); break; /* :116-118 */
        case OP_OFFSET: last.w = last.w - p[0]; break;    /* :124-126 */
        case OP_SHELL: /* :128-131 */
            if (!(last.w >= 0.0f)) last = neg(last);
            last.w = last.w - p[0];
            break;
        case OP_REPETITION: /* unsafe.cl:1-6 */
            last = mk(cc_remainder(last.x, p[0]), cc_remainder(last.y, p[1]),
                      cc_remainder(last.z, p[2]), 0);
            break;
        case OP_CIRCULAR_REPETITION_TO: last = circular_repetition_to(I, last); break;
        case OP_CIRCULAR_REPETITION_FROM: last = circular_repetition_from(I, last, in2); break;
        case OP_INVOLUTE_GEAR: last = involute_gear(I, last); break;
        case OP_EXTRUSION: /* simple3d.cl:18-21 */
            last = perpendicular_intersection(slab_z(p[0], in2), last);
            break;
        case OP_REVOLUTION_FROM: { /* simple3d.cl:28-39 */
            float cx = in2.x, cz = in2.z;
            float len = cc_len2(cx, cz), mult;
            if (len == 0.0f) { cx = 1.0f; mult = last.x; }
            else mult = cc_div(last.x, len);
            last = mk(cx * mult, last.y, cz * mult, last.w);
            break;
        }
        case OP_TWIST_REVOLUTION_FROM: last = twist_revolution_from(I, last, in2); break;
        case OP_SYMMETRICAL_FROM: /* common.cl:120-122 */
            if (in2.x < 0.0f) last.x = -last.x;
            break;
        case OP_UNION: last = rounded_union(p[0], last, in2); break;              /* :66-68 */
        case OP_INTERSECTION: last = neg(rounded_union(p[0], neg(last), neg(in2))); break; /* :70-72 */
        case OP_SUBTRACTION: last = neg(rounded_union(p[0], neg(last), in2)); break;       /* :74-76 */
        default: return mk(NAN, NAN, NAN, NAN);
        }
    }
#undef in2
}

static inline float grid_coord(float corner, float step, unsigned i)
{
    /* grid_eval.cl:13,31  corner + step * convert_float(i), contracted to one FMA */
    return cc_fma(step, (float)i, corner);
}

/* ---- exported entry points -------------------------------------------------------- */

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* returns the number of instructions, <0 on a malformed program */
int oracle_validate(const float *words, int n_words)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
    rc = prog.count;
    release(&prog);
    return rc;
}

/* evaluate at arbitrary points: pts[n][3] -> out[n][4] */
int oracle_evaluate_points(const float *words, int n_words, const float *pts, long n, float *out)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        v4 r = evaluate(&prog, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        memcpy(out + 4 * i, &r, sizeof(r));
    }
    release(&prog);
    return 0;
}

/* grid_eval.cl:23-34; the x index of the launch is x_offset + [0, nx) so that a
 * slab of a larger grid evaluates bit-identically to the unsharded launch. */
int oracle_grid_eval(const float *words, int n_words, const float *corner, float step,
                     int nx, int ny, int nz, int x_offset, float *out)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            float px = grid_coord(corner[0], step, (unsigned)(x + x_offset));
            float py = grid_coord(corner[1], step, (unsigned)y);
            for (int z = 0; z < nz; ++z) {
                v4 r = evaluate(&prog, px, py, grid_coord(corner[2], step, (unsigned)z));
                size_t idx = (size_t)z + (size_t)nz * ((size_t)y + (size_t)ny * (size_t)x);
                memcpy(out + 4 * idx, &r, sizeof(r));
            }
        }
    release(&prog);
    return 0;
}

/* Mean executed-branch algorithmic flop/point over every `stride`-th point per axis of the grid
 * (a stratified subsample; SURVEY.md 8(d): 64^3 of the config grid). */
int oracle_executed_flops(const float *words, int n_words, const float *corner, float step,
                          int nx, int ny, int nz, int x_offset, int stride, double *mean, long long *points)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
    if (stride < 1) stride = 1;
    unsigned long long total = 0;
    long long count = 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(+ : total, count)
    for (int x = stride / 2; x < nx; x += stride)
        for (int y = stride / 2; y < ny; y += stride) {
            float px = grid_coord(corner[0], step, (unsigned)(x + x_offset));
            float py = grid_coord(corner[1], step, (unsigned)y);
            for (int z = stride / 2; z < nz; z += stride) {
                t_flops = 0;
                (void)evaluate(&prog, px, py, grid_coord(corner[2], step, (unsigned)z));
                total += t_flops;
                ++count;
            }
        }
    release(&prog);
    *mean = count ? (double)total / (double)count : 0.0;
    *points = count;
    return 0;
}

/* grid_eval.cl:2-21 (distance only, y-flipped PyMCubes layout) */
int oracle_grid_eval_pymcubes(const float *words, int n_words, const float *corner, float step,
                              int nx, int ny, int nz, float *out)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            float px = grid_coord(corner[0], step, (unsigned)x);
            float py = grid_coord(corner[1], step, (unsigned)y);
            for (int z = 0; z < nz; ++z) {
                v4 r = evaluate(&prog, px, py, grid_coord(corner[2], step, (unsigned)z));
                size_t idx = (size_t)z + ((size_t)x + (size_t)(ny - y - 1) * (size_t)nx) * (size_t)nz;
                out[idx] = r.w;
            }
        }
    release(&prog);
    return 0;
}

/* subdivision.cl:12-30.  The reference appends with atomic_inc (arbitrary order);
 * the oracle emits in INDEX3 order (x slowest, z fastest), which is also the order
 * of the device's prefix-sum compaction. list = uchar4 (x, y, z, 0). */
int oracle_subdivision_step(const float *words, int n_words, const float *corner, float step,
                            float threshold, int nx, int ny, int nz, uint32_t *counter,
                            uint8_t *list)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
    size_t total = (size_t)nx * ny * nz;
    uint8_t *flag = (uint8_t *)malloc(total);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            float px = grid_coord(corner[0], step, (unsigned)x);
            float py = grid_coord(corner[1], step, (unsigned)y);
            for (int z = 0; z < nz; ++z) {
                float v = evaluate(&prog, px, py, grid_coord(corner[2], step, (unsigned)z)).w;
                flag[(size_t)z + (size_t)nz * ((size_t)y + (size_t)ny * (size_t)x)] =
                    (v > -threshold && v < threshold);
            }
        }
    uint32_t c = *counter; /* list[atomic_inc(counter)]: append after existing entries */
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z)
                if (flag[(size_t)z + (size_t)nz * ((size_t)y + (size_t)ny * (size_t)x)]) {
                    list[4 * c + 0] = (uint8_t)x; list[4 * c + 1] = (uint8_t)y;
                    list[4 * c + 2] = (uint8_t)z; list[4 * c + 3] = 0;
                    ++c;
                }
    *counter = c;
    free(flag);
    release(&prog);
    return 0;
}

/* mass_properties.cl:7-56.  sums (uint32, wrap-around like the device) in the
 * reference's order xx,xy,xz,x,yy,yz,y,zz,z,n (:34-41). */
int oracle_mass_properties_step(const float *words, int n_words, const float *corner, float step,
                                float threshold, int nx, int ny, int nz, uint32_t *sums,
                                uint32_t *counter, uint8_t *list)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
    size_t total = (size_t)nx * ny * nz;
    uint8_t *flag = (uint8_t *)malloc(total);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            float px = grid_coord(corner[0], step, (unsigned)x);
            float py = grid_coord(corner[1], step, (unsigned)y);
            for (int z = 0; z < nz; ++z) {
                float v = evaluate(&prog, px, py, grid_coord(corner[2], step, (unsigned)z)).w;
                uint8_t f = 0;
                if (v <= -threshold) f = 1;
                else if (v < threshold) f = 2;
                flag[(size_t)z + (size_t)nz * ((size_t)y + (size_t)ny * (size_t)x)] = f;
            }
        }
    uint32_t c = *counter, s[10] = {0};
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                uint8_t f = flag[(size_t)z + (size_t)nz * ((size_t)y + (size_t)ny * (size_t)x)];
                if (f == 1) {
                    uint32_t co[4] = {(uint32_t)x, (uint32_t)y, (uint32_t)z, 1u};
                    int i = 0;
                    for (int j = 0; j < 4; ++j)
                        for (int k = j; k < 4; ++k) s[i++] += co[j] * co[k];
                } else if (f == 2) {
                    list[4 * c + 0] = (uint8_t)x; list[4 * c + 1] = (uint8_t)y;
                    list[4 * c + 2] = (uint8_t)z; list[4 * c + 3] = 0;
                    ++c;
                }
            }
    for (int i = 0; i < 10; ++i) sums[i] += s[i];
    *counter = c;
    free(flag);
    release(&prog);
    return 0;
}

/* scalar access to the canonical math layer, for tests/test_oracle_math.py */
void oracle_math_probe(int which, const float *a, const float *b, long n, float *out, float *out2)
{
    for (long i = 0; i < n; ++i) {
        switch (which) {
        case 0: out[i] = cc_atan2(a[i], b[i]); break;
        case 1: cc_sincos(a[i], &out[i], &out2[i]); break;
        case 2: out[i] = cc_acos(a[i]); break;
        case 3: out[i] = cc_fmod_pos(a[i], b[i]); break;
        case 4: out[i] = cc_remainder(a[i], b[i]); break;
        default: out[i] = NAN;
        }
    }
}

/* ---- image renderers: restated ONLY to pin evaluate() against the reference's golden PNGs ----------
 * (tests/baseline/rendered_*.png, tests/test_image.py:16-28: the one place where real outputs of the
 * reference's OpenCL evaluate() are stored, for all 32 shapes of tests/data.py).  Not part of the
 * product: the ray caster is out of scope. */

typedef struct { float x, y, z; } v3;
static inline v3 v3mk(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3add(v3 a, v3 b) { return v3mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3scale(v3 a, float s) { return v3mk(a.x * s, a.y * s, a.z * s); }
static inline float v3dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3normalize(v3 a) { float l = sqrtf(v3dot(a, a)); return v3mk(a.x / l, a.y / l, a.z / l); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline v4 eval_at(const prog_t *prog, v3 p) { return evaluate(prog, p.x, p.y, p.z); }

/* ray_caster.cl:14-26 */
static float over_relaxation_step(v3 direction, v4 r)
{
    float over = 0.5f * fminf(1.0f, 1.0f + v3dot(direction, v3mk(r.x, r.y, r.z)));
    return r.w * (1 + over);
}
/* ray_caster.cl:28-40 */
static void light_no_trace(v3 normal, v3 toLight, v3 toCamera, float *diffuse, float *specular)
{
    v3 halfway = v3normalize(v3add(toLight, toCamera));
    *diffuse = fmaxf(0.0f, v3dot(normal, toLight));
    float sp = fmaxf(0.0f, v3dot(normal, halfway));
    sp *= sp; sp *= sp; sp *= sp;
    *specular = sp;
}
/* ray_caster.cl:42-99; returns the step count (what the false-colour mode reports) */
static unsigned light_contribution(const prog_t *prog, v3 point, v3 normal, v3 toLight, v3 toCamera, float minDistance,
                                   float maxDistance, float *diffuse, float *specular)
{
    light_no_trace(normal, toLight, toCamera, diffuse, specular);
    if (*diffuse <= 0 && *specular <= 0) { *diffuse = *specular = 0; return 0; }
    float threshold = (1.0f / 128.0f) / fmaxf(*diffuse, *specular);
    float visibility = 1, distance = minDistance, fallback = distance;
    unsigned step;
    for (step = 0; step < 100; ++step) {
        v4 r = eval_at(prog, v3add(point, v3scale(toLight, distance)));
        visibility = fminf(visibility, r.w / distance);
        if (visibility < threshold) break;
        if (distance - fallback > r.w) { distance = fallback; continue; }
        fallback = distance + r.w;
        distance = distance + over_relaxation_step(toLight, r);
        if (distance > maxDistance) break;
    }
    *diffuse *= visibility;
    *specular *= visibility;
    return step;
}
/* ray_caster.cl:101-117 */
static float ambient_occlusion(const prog_t *prog, v3 point, v3 normal, float distanceStep)
{
    float occlusion = 0.0f, scale = 1.0f, distance = distanceStep;
    for (unsigned i = 0; i < 4; ++i) {
        v4 r = eval_at(prog, v3add(point, v3scale(normal, distance)));
        occlusion += scale * (distance - r.w);
        scale /= 2;
        distance += distanceStep;
    }
    return clampf(1 - occlusion * 0.5f / (1 - scale), 0.0f, 1.0f);
}
static inline float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3 - 2 * t);
}

/* ray_caster.cl:147-256, one pixel; out = 3 bytes */
static void ray_pixel(const prog_t *prog, v3 origin, v3 forward, v3 up, v3 right, float pixelTolerance, float boxRadius,
                      float minDistance, float maxDistance, float floorZ, unsigned renderOptions, int x, int y, int w, int h,
                      unsigned char *out)
{
    float filmx = x - (w - 1) / 2.0f, filmy = y - (h - 1) / 2.0f;
    v3 direction = v3normalize(v3add(v3add(forward, v3scale(right, filmx)), v3scale(up, -filmy)));
    float distance = minDistance, fallback = minDistance;
    v4 r = {0, 0, 0, 0};
    int hit = 0;
    unsigned step;
    for (step = 0; step < 1000; ++step) {
        r = eval_at(prog, v3add(origin, v3scale(direction, distance)));
        if (distance - fallback > r.w) { distance = fallback; continue; }
        hit = r.w < pixelTolerance * distance;
        if (hit) {
            distance += r.w * clampf(1.0f / v3dot(v3mk(r.x, r.y, r.z), v3scale(direction, -1.0f)), 0.0f, 2.0f);
            break;
        }
        fallback = distance + r.w;
        distance = distance + over_relaxation_step(direction, r);
        if (distance > maxDistance) { distance = INFINITY; break; }
    }
    float cr, cg, cb;
    float localEpsilon = fmaxf(1e-4f, 2 * fabsf(r.w));
    const v3 light_dir = v3normalize(v3mk(1, 2, -1)), light2_dir = v3normalize(v3mk(-1, 1, 0));
    if (renderOptions & 1u) { /* RENDER_OPTIONS_FALSE_COLOR, ray_caster.cl:203-221 */
        v3 point = v3add(origin, v3scale(direction, distance));
        v3 normal = v3mk(r.x, r.y, r.z);
        float residual = hit ? fabsf(eval_at(prog, point).w) : 0.0f;
        float steps = (float)step, d1, s1;
        steps += (float)light_contribution(prog, point, normal, v3scale(light_dir, -1.0f), v3scale(direction, -1.0f),
                                           localEpsilon, maxDistance, &d1, &s1);
        steps += 4;
        cr = steps; cg = 1000 * residual; cb = 0;
    } else if (hit) {
        v3 point = v3add(origin, v3scale(direction, distance));
        v3 normal = v3mk(r.x, r.y, r.z);
        float ambient = ambient_occlusion(prog, point, normal, boxRadius / 100);
        float d1, s1, d2, s2;
        light_contribution(prog, point, normal, v3scale(light_dir, -1.0f), v3scale(direction, -1.0f), localEpsilon, maxDistance,
                           &d1, &s1);
        light_no_trace(normal, v3scale(light2_dir, -1.0f), v3scale(direction, -1.0f), &d2, &s2);
        float diffuse = 0.8f * d1 + 0.2f * d2, specular = 0.8f * s1 + 0.2f * s2;
        if (renderOptions & 2u) { /* RENDER_OPTIONS_ZEBRA: map_color_zebra, ray_caster.cl:134-145 */
            int white = (int)floorf(point.y) & 1;
            float color = 50 + 150 * white;
            color *= ambient + diffuse;
            color += 128 * specular;
            cr = cg = cb = color;
        } else {
            /* map_color, ray_caster.cl:119-132 */
            float saturation = 0.75f * smoothstepf(0.0f, 0.25f, diffuse);
            float value = 0.1f + 0.8f * (diffuse + (ambient - diffuse) * 0.3f);
            float chroma = value * saturation, X = chroma * 0.7f, m = value - chroma;
            cr = 255 * (X + m) + specular * 128;
            cg = 255 * (chroma + m) + specular * 128;
            cb = 255 * m + specular * 128;
        }
    } else {
        cr = 230; cg = 230; cb = 241;
    }
    float floorDistance = (floorZ - origin.z) / direction.z;
    if (floorDistance > 0 && floorDistance < distance) {
        v3 floorPoint = v3add(origin, v3scale(direction, floorDistance));
        float fd = eval_at(prog, floorPoint).w;
        float shadow = clampf(2 * fd / boxRadius, 0.0f, 1.0f);
        shadow = 1 - shadow; shadow *= shadow; shadow = 1 - shadow;
        float k = 0.4f + 0.6f * shadow;
        cr *= k; cg *= k; cb *= k;
    }
    out[0] = (unsigned char)clampf(cr, 0.0f, 255.0f);
    out[1] = (unsigned char)clampf(cg, 0.0f, 255.0f);
    out[2] = (unsigned char)clampf(cb, 0.0f, 255.0f);
}

/* out[x][y][3] with INDEX2 (y fastest), like the reference's output buffer (ray_caster.py:49) */
int oracle_ray_caster(const float *words, int n_words, const float *origin, const float *forward, const float *up,
                      const float *right, float pixelTolerance, float boxRadius, float minDistance, float maxDistance,
                      float floorZ, unsigned renderOptions, int w, int h, unsigned char *out)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
    v3 o = v3mk(origin[0], origin[1], origin[2]), f = v3mk(forward[0], forward[1], forward[2]);
    v3 u = v3mk(up[0], up[1], up[2]), r = v3mk(right[0], right[1], right[2]);
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y)
            ray_pixel(&prog, o, f, u, r, pixelTolerance, boxRadius, minDistance, maxDistance, floorZ, renderOptions, x, y, w, h,
                      out + 3 * ((size_t)y + (size_t)h * (size_t)x));
    release(&prog);
    return 0;
}

/* bitmap.cl:1-18 */
int oracle_bitmap(const float *words, int n_words, const float *origin, float stepSize, int w, int h, unsigned char *out)
{
    prog_t prog;
    int rc = prepare(words, n_words, &prog);
    if (rc < 0) return rc;
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            float d = evaluate(&prog, origin[0] + stepSize * (float)x, origin[1] + stepSize * (float)(h - y - 1), origin[2]).w;
            unsigned char *px = out + 3 * ((size_t)y + (size_t)h * (size_t)x);
            if (d < 0.0f) { px[0] = 125; px[1] = 179; px[2] = 0; }   /* mix(inside, background, step(0, d)) */
            else { px[0] = 230; px[1] = 230; px[2] = 241; }
        }
    release(&prog);
    return 0;
}

/* ---- 2-D outline extraction: rendering/polygon2d.cl ------------------------------------------------
 * process_polygon over global size (cx, cy, 2): `corners` is the float4 grid [cx+1][cy+1] that grid_eval
 * produced (INDEX2 with sizes cx+1, cy+1).  Canonical arithmetic: single IEEE operations in source
 * order.  `starts` is filled in increasing cell-index order (the reference's atomic_inc order is
 * arbitrary). */

/* polygon2d.cl:5-35 */
static uint32_t poly_encode_index(int cx, int cy, uint32_t index, int gs0, int gs1)
{
    index &= (1u << 20) - 1u;
    int is_y, out_c, other_c;
    if (cx < 0 || cx >= gs0) { is_y = 0; out_c = cx; other_c = cy; }
    else if (cy < 0 || cy >= gs1) { is_y = 1; out_c = cy; other_c = cx; }
    else return index;
    return 0x80000000u | (is_y ? 0x40000000u : 0u) | (out_c < 0 ? 0x20000000u : 0u) | ((uint32_t)other_c << 20) | index;
}

/* polygon2d.cl:37-80 */
static void poly_place_vertex(const float pos[3][2], const v4 val[3], float out[2])
{
    float ax = 0, ay = 0, weight = 0;
    for (int i = 0; i < 3; ++i) {
        float w = 1 / (1 + fabsf(val[i].w));
        ax += pos[i][0] * w;
        ay += pos[i][1] * w;
        weight += w;
    }
    ax /= weight;
    ay /= weight;
    float px = ax, py = ay;
    for (int i = 0; i < 8; ++i) {
        float gx = 0, gy = 0, residualSum = 0;
        for (int j = 0; j < 3; ++j) {
            float nx = val[j].x, ny = val[j].y;
            float tmp = (nx * (px - pos[j][0]) + ny * (py - pos[j][1])) + val[j].w;
            residualSum += tmp * tmp;
            gx += nx * tmp;
            gy += ny * tmp;
        }
        if (residualSum < 1e-3f) break;
        float gl = gx * gx + gy * gy;
        if (gl < 1e-8f) break;
        float k = residualSum / gl;
        px -= gx * k;
        py -= gy * k;
    }
    out[0] = px;
    out[1] = py;
}

/* polygon2d.cl:82-175.  vertices: float2 per cell (only written for surface cells), links: uint32 per cell */
int oracle_process_polygon(const float *box_corner, float box_step, int gs0, int gs1, const float *corners, float *vertices,
                           uint32_t *links, uint32_t *starts, uint32_t *start_counter)
{
    const v4 *c4 = (const v4 *)corners;
    uint32_t n_starts = 0;
    for (int x = 0; x < gs0; ++x)
        for (int y = 0; y < gs1; ++y)
            for (int t = 0; t < 2; ++t) {
                const int off[3][2] = {{0, 0}, {1, 1}, {t, 1 - t}};
                unsigned cellType = 0;
                for (int i = 0; i < 3; ++i) {
                    size_t ci = (size_t)(y + off[i][1]) + (size_t)(gs1 + 1) * (size_t)(x + off[i][0]);
                    cellType = cellType << 1 | (c4[ci].w <= 0 ? 1u : 0u);
                }
                const uint32_t index = (uint32_t)t + 2u * ((uint32_t)y + (uint32_t)gs1 * (uint32_t)x);
                if (cellType == 0 || cellType == 7) { links[index] = 0xFFFFFFFFu; continue; }
                int backwards = cellType == 3 || cellType == 5 || cellType == 6;
                if (backwards) cellType = 7 - cellType;
                const int flip = t == 1;
                if (flip) backwards = !backwards;
                int fx = 0, fy = 0, rx = 0, ry = 0;
                switch (cellType) {
                case 1: fx = 0; fy = 1; rx = -1; ry = 0; break;
                case 2: fx = 0; fy = 0; rx = 0; ry = 1; break;
                case 4: fx = -1; fy = 0; rx = 0; ry = 0; break;
                }
                if (backwards) { int a = fx, b = fy; fx = rx; fy = ry; rx = a; ry = b; }
                if (flip) { int a = fx; fx = fy; fy = a; a = rx; rx = ry; ry = a; }
                fx += x; fy += y; rx += x; ry += y;
                links[index] = poly_encode_index(fx, fy, (uint32_t)(1 - t) + 2u * ((uint32_t)fy + (uint32_t)gs1 * (uint32_t)fx),
                                                 gs0, gs1);
                uint32_t startIndex = poly_encode_index(rx, ry, index, gs0, gs1);
                if (startIndex & 0x80000000u) starts[n_starts++] = startIndex ^ 0x20000000u;
                float pos[3][2];
                v4 val[3];
                for (int i = 0; i < 3; ++i) {
                    int cx = x + off[i][0], cy = y + off[i][1];
                    pos[i][0] = box_corner[0] + (float)cx * box_step;
                    pos[i][1] = box_corner[1] + (float)cy * box_step;
                    val[i] = c4[(size_t)cy + (size_t)(gs1 + 1) * (size_t)cx];
                }
                poly_place_vertex((const float(*)[2])pos, val, vertices + 2 * (size_t)index);
            }
    *start_counter = n_starts;
    return 0;
}
