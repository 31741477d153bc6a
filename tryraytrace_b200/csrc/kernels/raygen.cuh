// raygen.cuh -- primary-ray generation, bit-equal to the reference
// (reference src/renderer.cu:331-356 as compiled for sm_100: SURVEY Appendix A.2,
// re-checked against the PTX/SASS of the unmodified kernel).
#pragma once
#include "common.cuh"
#include "xorwow.cuh"

namespace trt {

TRT_DEV float p_cos(float a) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
TRT_DEV float p_sin(float a) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// tent filter offset in [-1,1] from r = 2*u  (:333-334)
TRT_DEV float tent(float r) {
    return r < 1.f ? p_add(p_sqrt(r), -1.f) : p_sub(1.f, p_sqrt(p_sub(2.f, r)));
}

// float(double(u) * 2*pi): the reference multiplies by the double constant M_PI
TRT_DEV float two_pi_times(float u) { return (float)((double)u * 6.283185307179586); }

// x in [0,w), y in [0,h) with y pointing up; the RNG must be positioned at the start of the
// pixel's stream.  Draw order: jitter x, jitter y, then (lens only) radius, angle.
TRT_DEV Ray primary_ray(const Camera& cam, int x, int y, int w, int h, Xorwow& rng) {
    const float u0 = xw_uniform(rng);
    const float u1 = xw_uniform(rng);
    const float dx = tent(p_add(u0, u0));
    const float dy = tent(p_add(u1, u1));
    // (x + .5 + dx) / w - .5: ptxas expands the approximate division into rcp * numerator and
    // contracts that multiply with the following subtraction (FFMA num, rcp(w), -0.5 in the
    // reference SASS), so the quotient is never rounded on its own
    const float sx = p_fma(p_add(p_add((float)x, 0.5f), dx), p_rcp((float)w), -0.5f);
    const float sy = p_fma(p_add(p_add((float)y, 0.5f), dy), p_rcp((float)h), -0.5f);
    F3 dir = f3(p_add(cam.dir.x, p_fma(cam.cx.x, sx, p_mul(cam.cy.x, sy))),
                p_add(cam.dir.y, p_fma(cam.cx.y, sx, p_mul(cam.cy.y, sy))),
                p_add(cam.dir.z, p_fma(cam.cx.z, sx, p_mul(cam.cy.z, sy))));
    dir = x_normalize(dir);

    F3 lens = f3(0.f, 0.f, 0.f);
    if (cam.lens_radius > 0.f) {
        const float lr = p_mul(cam.lens_radius, p_sqrt(xw_uniform(rng)));
        const float th = two_pi_times(xw_uniform(rng));
        const F3 u = x_normalize(f3(cam.cx.x, cam.cx.y, cam.cx.z));
        const F3 v = x_normalize(f3(cam.cy.x, cam.cy.y, cam.cy.z));
        const float a = p_mul(lr, p_cos(th));
        const float b = p_mul(lr, p_sin(th));
        lens = f3(p_fma(u.x, a, p_mul(v.x, b)), p_fma(u.y, a, p_mul(v.y, b)), p_fma(u.z, a, p_mul(v.z, b)));
    }
    const F3 focus = f3(p_fma(cam.focus_dist, dir.x, cam.pos.x), p_fma(cam.focus_dist, dir.y, cam.pos.y),
                        p_fma(cam.focus_dist, dir.z, cam.pos.z));
    Ray r;
    r.o = f3(p_add(cam.pos.x, lens.x), p_add(cam.pos.y, lens.y), p_add(cam.pos.z, lens.z));
    r.d = x_normalize(x_sub(focus, r.o));
    return r;
}

}  // namespace trt
