// Image renderers around evaluate(), independent of HOW a point is evaluated (shared by the
// interpreter kernels in cc_kernels.cu and the NVRTC scene-specialised kernels of cc_jit.cpp).
// 
//
//   cc_bitmap_body      rendering/bitmap.cl:1-18        (2-D shapes: inside / background colour)
//   cc_ray_caster_body  rendering/ray_caster.cl:147-256 (3-D shapes: enhanced sphere tracing with
//                       over-relaxation, soft shadow ray, 4-tap ambient occlusion, floor shadow)
//
// Both run the SAME evaluate() as the grid kernels, one point per thread.  The op library votes
// across the warp, so evaluate() must be reached by all 32 lanes together; a ray caster whose
// rays each loop a different number of times therefore cannot call it from per-thread control flow.
// The kernel is written as a per-lane state machine around ONE warp-convergent evaluate() per
// round: every lane proposes the next point of whatever it is doing (primary ray, residual probe,
// shadow ray, ambient-occlusion tap, floor probe), the warp evaluates, every lane consumes its own
// result and advances.  Lanes in different phases share rounds, so a warp is done after
// max-over-lanes(total evaluations) rounds instead of the sum of per-phase maxima, and the
// evaluation code is instantiated once.  A warp owns an 8x4 pixel tile (coherent rays -> similar
// round counts).  Arithmetic: plain IEEE single operations in source order (-fmad=false; div.rn,
// sqrt.rn), the same sequence as the CPU oracle's restatement, so the images are bit-identical.

#ifndef CC_RENDER_CUH
#define CC_RENDER_CUH

#include "cc_device_types.h"
#include "cc_math.cuh"


struct cc_f3 {
    float x, y, z;
};
CC_DEV cc_f3 f3(float x, float y, float z) { cc_f3 r = {x, y, z}; return r; }
CC_DEV cc_f3 f3add(cc_f3 a, cc_f3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
CC_DEV cc_f3 f3scale(cc_f3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
CC_DEV float f3dot(cc_f3 a, cc_f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CC_DEV cc_f3 f3normalize(cc_f3 a)
{
    const float l = __fsqrt_rn(f3dot(a, a));
    return f3(__fdiv_rn(a.x, l), __fdiv_rn(a.y, l), __fdiv_rn(a.z, l));
}
CC_DEV float cc_clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// ray_caster.cl:14-26
CC_DEV float cc_over_relaxation_step(cc_f3 direction, float4 r)
{
    const float over = 0.5f * fminf(1.0f, 1.0f + f3dot(direction, f3(r.x, r.y, r.z)));
    return r.w * (1 + over);
}
// ray_caster.cl:28-40
CC_DEV void cc_light_no_trace(cc_f3 normal, cc_f3 toLight, cc_f3 toCamera, float *diffuse, float *specular)
{
    const cc_f3 halfway = f3normalize(f3add(toLight, toCamera));
    *diffuse = fmaxf(0.0f, f3dot(normal, toLight));
    float sp = fmaxf(0.0f, f3dot(normal, halfway));
    sp *= sp; sp *= sp; sp *= sp;
    *specular = sp;
}

// EVAL: functor  void operator()(const float (&gx)[1], const float (&gy)[1], const float (&gz)[1], cc_val<float> (&L)[1])
template <class EVAL>
CC_DEV float4 cc_eval_point(EVAL &eval, cc_f3 p)
{
    float gx[1] = {p.x}, gy[1] = {p.y}, gz[1] = {p.z};
    cc_val<float> L[1];
    eval(gx, gy, gz, L);
    return make_float4(L[0].x, L[0].y, L[0].z, L[0].w);
}

// pixel of this lane: warps tile the image in 8 (x) by 4 (y) pixel patches
CC_DEV bool cc_render_pixel(const cc_render_args &a, uint32_t *x, uint32_t *y)
{
    const uint32_t warp = (blockIdx.x * CC_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_y = (a.h + 3) / 4;
    const uint32_t tx = warp / tiles_y, ty = warp - tx * tiles_y;
    *x = tx * 8 + (lane >> 2);
    *y = ty * 4 + (lane & 3);
    return *x < a.w && *y < a.h;
}

CC_DEV void cc_store_pixel(const cc_render_args &a, uint32_t x, uint32_t y, float cr, float cg, float cb)
{
    uint8_t *px = a.out + 3 * ((size_t)y + (size_t)a.h * (size_t)x);
    px[0] = (uint8_t)cc_clampf(cr, 0.0f, 255.0f);
    px[1] = (uint8_t)cc_clampf(cg, 0.0f, 255.0f);
    px[2] = (uint8_t)cc_clampf(cb, 0.0f, 255.0f);
}

template <class EVAL>
CC_DEV void cc_bitmap_body(const cc_render_args &a, EVAL &eval)
{
    uint32_t x, y;
    const bool valid = cc_render_pixel(a, &x, &y);
    // bitmap.cl:5-8
    const cc_f3 p = f3(a.ox + a.step_size * (float)x, a.oy + a.step_size * (float)(a.h - y - 1), a.oz);
    const float d = cc_eval_point(eval, valid ? p : f3(a.ox, a.oy, a.oz)).w;
    if (!valid) return;
    // mix(inside, background, step(0, d)): step(edge, x) = x < edge ? 0 : 1 (NaN -> 1)
    if (d < 0.0f) cc_store_pixel(a, x, y, 125.f, 179.f, 0.f);
    else cc_store_pixel(a, x, y, 230.f, 230.f, 241.f);
}

enum { RC_PRIMARY = 0, RC_RESIDUAL, RC_AO, RC_SHADOW, RC_FLOOR, RC_DONE };

template <class EVAL>
CC_DEV void cc_ray_caster_body(const cc_render_args &a, EVAL &eval)
{
    uint32_t x, y;
    const bool valid = cc_render_pixel(a, &x, &y);
    const bool false_color = (a.options & 1u) != 0, zebra = (a.options & 2u) != 0;
    const cc_f3 origin = f3(a.ox, a.oy, a.oz);

    // ray_caster.cl:160-165
    const float filmx = (float)x - (float)(a.w - 1) / 2.0f, filmy = (float)y - (float)(a.h - 1) / 2.0f;
    const cc_f3 direction = f3normalize(f3add(f3add(f3(a.fx, a.fy, a.fz), f3scale(f3(a.rx, a.ry, a.rz), filmx)),
                                              f3scale(f3(a.ux, a.uy, a.uz), -filmy)));
    const cc_f3 to_camera = f3scale(direction, -1.0f);
    const cc_f3 to_light = f3scale(f3normalize(f3(1, 2, -1)), -1.0f);
    const cc_f3 to_light2 = f3scale(f3normalize(f3(-1, 1, 0)), -1.0f);

    int phase = valid ? RC_PRIMARY : RC_DONE;
    float distance = a.min_distance, fallback = a.min_distance;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    bool hit = false;
    uint32_t step = 0;
    // state of the secondary phases
    cc_f3 point = origin, normal = f3(0.f, 0.f, 0.f);
    float local_epsilon = 0.f, residual = 0.f;
    float occlusion = 0.f, ao_scale = 1.f, ao_distance = 0.f, ambient = 0.f;
    const float ao_step = a.box_radius / 100;
    uint32_t ao_i = 0;
    float d1 = 0.f, s1 = 0.f, threshold = 0.f, visibility = 1.f, sdist = 0.f, sfallback = 0.f;
    uint32_t sstep = 0;
    float cr = 0.f, cg = 0.f, cb = 0.f;
    float floor_distance = 0.f;
    uint32_t evals = 0;

    for (;;) {
        if (!__any_sync(0xffffffffu, phase != RC_DONE)) break;
        cc_f3 q;
        switch (phase) {
        case RC_PRIMARY: q = f3add(origin, f3scale(direction, distance)); break;
        case RC_RESIDUAL: q = point; break;
        case RC_AO: q = f3add(point, f3scale(normal, ao_distance)); break;
        case RC_SHADOW: q = f3add(point, f3scale(to_light, sdist)); break;
        case RC_FLOOR: q = f3add(origin, f3scale(direction, floor_distance)); break;
        default: q = origin; break;
        }
        __syncwarp();
        const float4 e = cc_eval_point(eval, q);
        if (phase == RC_DONE) continue;
        ++evals;

        // what the lane does once the current phase has produced its last value
        enum { GO_NONE = 0, GO_AFTER_PRIMARY, GO_LIGHT, GO_SHADE, GO_FLOOR } go = GO_NONE;

        if (phase == RC_PRIMARY) {
            // ray_caster.cl:172-199
            r = e;
            bool done = false;
            if (distance - fallback > r.w) {
                distance = fallback;  // over-relaxation was too optimistic
            } else {
                hit = r.w < a.pixel_tolerance * distance;
                if (hit) {
                    distance += r.w * cc_clampf(__fdiv_rn(1.0f, f3dot(f3(r.x, r.y, r.z), to_camera)), 0.0f, 2.0f);
                    done = true;
                } else {
                    fallback = distance + r.w;
                    distance = distance + cc_over_relaxation_step(direction, r);
                    if (distance > a.max_distance) {
                        distance = __int_as_float(0x7f800000);
                        done = true;
                    }
                }
            }
            if (!done) {
                ++step;
                done = step >= 1000u;
            }
            if (done) go = GO_AFTER_PRIMARY;
        } else if (phase == RC_RESIDUAL) {
            residual = fabsf(e.w);  // ray_caster.cl:210
            go = GO_LIGHT;
        } else if (phase == RC_AO) {
            // ray_caster.cl:101-117
            occlusion += ao_scale * (ao_distance - e.w);
            ao_scale = __fdiv_rn(ao_scale, 2.f);
            ao_distance += ao_step;
            if (++ao_i == 4u) {
                ambient = cc_clampf(1 - __fdiv_rn(occlusion * 0.5f, 1 - ao_scale), 0.0f, 1.0f);
                go = GO_LIGHT;
            }
        } else if (phase == RC_SHADOW) {
            // ray_caster.cl:57-91
            bool done = false;
            visibility = fminf(visibility, __fdiv_rn(e.w, sdist));
            if (visibility < threshold) {
                done = true;
            } else if (sdist - sfallback > e.w) {
                sdist = sfallback;
            } else {
                sfallback = sdist + e.w;
                sdist = sdist + cc_over_relaxation_step(to_light, e);
                if (sdist > a.max_distance) done = true;
            }
            if (!done) {
                ++sstep;
                done = sstep >= 100u;
            }
            if (done) {
                d1 *= visibility;
                s1 *= visibility;
                go = GO_SHADE;
            }
        } else {  // RC_FLOOR, ray_caster.cl:243-249
            float shadow = cc_clampf(__fdiv_rn(2 * e.w, a.box_radius), 0.0f, 1.0f);
            shadow = 1 - shadow;
            shadow *= shadow;
            shadow = 1 - shadow;
            const float k = 0.4f + 0.6f * shadow;
            cr *= k; cg *= k; cb *= k;
            cc_store_pixel(a, x, y, cr, cg, cb);
            phase = RC_DONE;
        }

        if (go == GO_AFTER_PRIMARY) {
            local_epsilon = fmaxf(1e-4f, 2 * fabsf(r.w));
            point = f3add(origin, f3scale(direction, distance));
            normal = f3(r.x, r.y, r.z);
            if (false_color) {
                if (hit) phase = RC_RESIDUAL;
                else { residual = 0.f; go = GO_LIGHT; }
            } else if (hit) {
                occlusion = 0.f; ao_scale = 1.f; ao_distance = ao_step; ao_i = 0;
                phase = RC_AO;
            } else {
                cr = 230.f; cg = 230.f; cb = 241.f;
                go = GO_FLOOR;
            }
        }
        if (go == GO_LIGHT) {
            // light_contribution, ray_caster.cl:42-55
            cc_light_no_trace(normal, to_light, to_camera, &d1, &s1);
            sstep = 0;
            if (d1 <= 0 && s1 <= 0) {
                d1 = s1 = 0.f;
                go = GO_SHADE;
            } else {
                threshold = __fdiv_rn(1.0f / 128.0f, fmaxf(d1, s1));
                visibility = 1.f;
                sdist = local_epsilon;
                sfallback = sdist;
                phase = RC_SHADOW;
            }
        }
        if (go == GO_SHADE) {
            if (false_color) {
                // ray_caster.cl:215-220
                float steps = (float)step;
                steps += (float)sstep;
                steps += 4;
                cr = steps; cg = 1000 * residual; cb = 0.f;
            } else {
                float d2, s2;
                cc_light_no_trace(normal, to_light2, to_camera, &d2, &s2);
                const float diffuse = 0.8f * d1 + 0.2f * d2, specular = 0.8f * s1 + 0.2f * s2;
                if (zebra) {
                    // map_color_zebra, ray_caster.cl:134-145
                    const int white = (int)floorf(point.y) & 1;
                    float color = 50 + 150 * white;
                    color *= ambient + diffuse;
                    color += 128 * specular;
                    cr = cg = cb = color;
                } else {
                    // map_color, ray_caster.cl:119-132
                    const float t = cc_clampf(__fdiv_rn(diffuse - 0.0f, 0.25f - 0.0f), 0.0f, 1.0f);
                    const float saturation = 0.75f * (t * t * (3 - 2 * t));
                    const float value = 0.1f + 0.8f * (diffuse + (ambient - diffuse) * 0.3f);
                    const float chroma = value * saturation, X = chroma * 0.7f, m = value - chroma;
                    cr = 255 * (X + m) + specular * 128;
                    cg = 255 * (chroma + m) + specular * 128;
                    cb = 255 * m + specular * 128;
                }
            }
            go = GO_FLOOR;
        }
        if (go == GO_FLOOR) {
            // ray_caster.cl:240-243
            floor_distance = __fdiv_rn(a.floor_z - origin.z, direction.z);
            if (floor_distance > 0 && floor_distance < distance) {
                phase = RC_FLOOR;
            } else {
                cc_store_pixel(a, x, y, cr, cg, cb);
                phase = RC_DONE;
            }
        }
    }
    if (a.eval_count) {
        const uint32_t total = __reduce_add_sync(0xffffffffu, evals);
        if ((threadIdx.x & 31) == 0 && total) atomicAdd(a.eval_count, (unsigned long long)total);
    }
}

#endif
