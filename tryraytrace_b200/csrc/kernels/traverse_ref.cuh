// traverse_ref.cuh -- TRAVERSE_REF: the reference's BVH2, node set, visit order and
// arithmetic (reference src/renderer.cu:371-425 closest hit, :273-314 any hit,
// include/aabb.h:49-69 slab test, src/renderer.cu:235-268 triangle test).
//
// This path defines the parity contract: first-hit ids, d_min and the three visit
// counters must equal the reference kernel's bit for bit.  It is also what ambiguous rays
// of the fast path are replayed through (traverse_wide.cuh).  It reads the arrays in the
// layout init_scene_data received them in (48-byte nodes, 112-byte objects), with 128-bit
// loads; the stack lives in registers/local memory like the reference's.
#pragma once
#include "common.cuh"

namespace trt {

// safe_inv, reference src/renderer.cu:371-373
TRT_DEV float ref_safe_inv(float x) {
    return (fabsf(x) < 1e-8f) ? (x >= 0.f ? 1e20f : -1e20f) : p_rcp(x);
}

// AABB::hit with the reference's operation order: (plane - o) * inv as FADD then FMUL,
// min/max as compare+select with the "NaN returns b" behaviour of fmin/fmax_wrapper.
TRT_DEV bool ref_slab(const float4 bmin, const float4 bmax, const F3 o, const F3 inv, float t_min, float t_max,
                      float* entry = nullptr) {
    const float tx1 = p_mul(p_sub(bmin.x, o.x), inv.x);
    const float tx2 = p_mul(p_sub(bmax.x, o.x), inv.x);
    float lo = tx1 < tx2 ? tx1 : tx2;
    float hi = tx1 > tx2 ? tx1 : tx2;
    const float ty1 = p_mul(p_sub(bmin.y, o.y), inv.y);
    const float ty2 = p_mul(p_sub(bmax.y, o.y), inv.y);
    const float ylo = ty1 < ty2 ? ty1 : ty2;
    const float yhi = ty1 > ty2 ? ty1 : ty2;
    lo = lo > ylo ? lo : ylo;
    hi = hi < yhi ? hi : yhi;
    const float tz1 = p_mul(p_sub(bmin.z, o.z), inv.z);
    const float tz2 = p_mul(p_sub(bmax.z, o.z), inv.z);
    const float zlo = tz1 < tz2 ? tz1 : tz2;
    const float zhi = tz1 > tz2 ? tz1 : tz2;
    lo = lo > zlo ? lo : zlo;
    hi = hi < zhi ? hi : zhi;
    if (entry) *entry = lo;
    return hi >= lo && hi > t_min && lo < t_max;
}

// Moeller-Trumbore on (v0, e1, e2) with e1 = v1 - v0, e2 = v2 - v0 already rounded the way
// the reference rounds them (one FADD each).  Returns t, or 0 for a miss.
TRT_DEV float ref_tri_edges(const F3 v0, const F3 e1, const F3 e2, const F3 o, const F3 d) {
    const float eps = 1e-5f;
    const F3 h = x_cross(d, e2);
    const float a = x_dot(e1, h);
    if (a > -eps && a < eps) return 0.f;
    const float f = p_rcp(a);
    const F3 s = x_sub(o, v0);
    const float u = p_mul(f, x_dot(s, h));
    if (u < 0.f || u > 1.f) return 0.f;
    const F3 q = x_cross(s, e1);
    const float v = p_mul(f, x_dot(d, q));
    if (v < 0.f || p_add(u, v) > 1.f) return 0.f;
    const float t = p_mul(f, x_dot(e2, q));
    return t > eps ? t : 0.f;
}

TRT_DEV float ref_tri_object(const float4* __restrict__ objects, int idx, const F3 o, const F3 d) {
    const float4* p = objects + (size_t)idx * 7;
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    const F3 v0 = f3(a.x, a.y, a.z);
    const F3 e1 = x_sub(f3(b.x, b.y, b.z), v0);
    const F3 e2 = x_sub(f3(c.x, c.y, c.z), v0);
    return ref_tri_edges(v0, e1, e2, o, d);
}

struct VisitCounts {
    uint32_t fetched, entered, tris;
};

// Closest hit in reference order.  COUNT adds the three visit counters of SURVEY 7.3(2).
template <bool COUNT>
TRT_DEV int ref_closest(const SceneDev& sc, const Ray& r, float* t_out, VisitCounts* vc) {
    const F3 inv = f3(ref_safe_inv(r.d.x), ref_safe_inv(r.d.y), ref_safe_inv(r.d.z));
    float d_min = 1e20f;
    int id = -1;
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4* np = sc.ref_nodes + (size_t)ni * 3;
        const float4 bmin = __ldg(np), bmax = __ldg(np + 1);
        const int4 link = __ldg(reinterpret_cast<const int4*>(np + 2));
        if (COUNT) vc->fetched++;
        if (!ref_slab(bmin, bmax, r.o, inv, 0.f, d_min)) continue;
        if (COUNT) vc->entered++;
        if (link.w) {  // leaf: link.x = first primitive, link.y = count
            for (int k = 0; k < link.y; k++) {
                const int oi = link.x + k;
                if (COUNT) vc->tris++;
                const float t = ref_tri_object(sc.objects, oi, r.o, r.d);
                if (t > 0.f && t < d_min) {
                    d_min = t;
                    id = oi;
                }
            }
        } else {
            stack[sp++] = link.y;  // right pushed first, left popped first
            stack[sp++] = link.x;
        }
    }
    *t_out = d_min;
    return id;
}

// Any hit in reference order (trace_shadow).  The inverse direction is the raw reciprocal
// (can be +-inf), the box interval is (0.001, max_dist), a triangle occludes iff
// 0.001 < t < max_dist - 0.001.
template <bool COUNT>
TRT_DEV bool ref_shadow(const SceneDev& sc, const Ray& r, float max_dist, VisitCounts* vc) {
    const F3 inv = f3(p_rcp(r.d.x), p_rcp(r.d.y), p_rcp(r.d.z));
    const float t_hi = p_sub(max_dist, 0.001f);
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4* np = sc.ref_nodes + (size_t)ni * 3;
        const float4 bmin = __ldg(np), bmax = __ldg(np + 1);
        const int4 link = __ldg(reinterpret_cast<const int4*>(np + 2));
        if (COUNT) vc->fetched++;
        if (!ref_slab(bmin, bmax, r.o, inv, 0.001f, max_dist)) continue;
        if (COUNT) vc->entered++;
        if (link.w) {
            for (int k = 0; k < link.y; k++) {
                if (COUNT) vc->tris++;
                const float t = ref_tri_object(sc.objects, link.x + k, r.o, r.d);
                if (t > 0.001f && t < t_hi) return true;
            }
        } else {
            stack[sp++] = link.y;
            stack[sp++] = link.x;
        }
    }
    return false;
}

}  // namespace trt
