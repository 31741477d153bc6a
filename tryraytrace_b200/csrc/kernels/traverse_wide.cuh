// traverse_wide.cuh -- TRAVERSE_FAST: ordered traversal of the 4-wide BVH with an exact
// accept rule, so that the result equals the reference traversal's
// (reference src/renderer.cu:371-425 / :273-314) without visiting its ~80 nodes per ray.
//
// Why the result is the same (DESIGN.md "exactness argument"):
//  1. Superset.  Every wide box contains the reference leaf boxes below it, and the slab
//     test here is the reference's own formula -- (plane - o) * inv, a rounded subtract then
//     a rounded multiply, with the same o and inv.  Both roundings are monotone, so a wider
//     box can only give an earlier entry and a later exit: whenever the reference's test of
//     a leaf box passes, the tests of all wide boxes above that triangle pass too.
//  2. Same triangle arithmetic.  Candidates are tested with the reference's Moeller-Trumbore
//     operation sequence (traverse_ref.cuh, ref_tri_edges) on pre-subtracted edges that
//     carry the same single rounding.
//  3. Exact reachability.  The reference tests triangle k only if the slab tests of all its
//     ancestors pass.  Ancestor boxes contain the leaf box, so by (1) they pass whenever the
//     leaf box passes with entry < t_k; a candidate whose reference leaf box fails outright
//     (exit < entry or exit <= t_min) can never be reached by the reference and is dropped.
//  4. Ambiguity.  The only order-dependent case left is a candidate whose leaf-box entry is
//     not below its own hit distance (the reference culls against the running d_min, which
//     depends on visit order).  Such rays are flagged and re-run through TRAVERSE_REF.
//     Ties in t resolve to the lowest object index, as the reference's ascending leaf order does.
// Any-hit queries have a fixed interval, so (3) decides them exactly and no replay exists.
//
// Execution model.  The traversal is written as ROUNDS so that the 32 rays of a warp stay
// converged: in every round each lane performs at most one wide-node step (four child-box
// tests) and at most one triangle step.  Inner children go to a per-lane node stack, leaf
// children to a separate per-lane triangle stack, so a lane that has found a leaf keeps
// descending speculatively instead of idling while its neighbours test boxes.  The
// persistent kernels in wavefront.cu refill idle lanes from the ray queue between rounds
// (while-while with dynamic fetch, Aila & Laine 2009).
#pragma once
#include "common.cuh"
#include "traverse_ref.cuh"

namespace trt {

struct WideCounts {
    uint32_t nodes, tris;
};

constexpr int kWideEmptyRef = 0x7fffffff;
constexpr int kNodeStack = 48;  // >= 3 * wide depth + 1 (checked at upload)
constexpr int kTriStack = 16;
constexpr float kCullSlack = 1.0005f;

// reference-formula slab for one child of a wide node
TRT_DEV void wide_slab(float lox, float hix, float loy, float hiy, float loz, float hiz, const F3 o, const F3 inv,
                       float* t_near, float* t_far) {
    const float x1 = p_mul(p_sub(lox, o.x), inv.x), x2 = p_mul(p_sub(hix, o.x), inv.x);
    const float y1 = p_mul(p_sub(loy, o.y), inv.y), y2 = p_mul(p_sub(hiy, o.y), inv.y);
    const float z1 = p_mul(p_sub(loz, o.z), inv.z), z2 = p_mul(p_sub(hiz, o.z), inv.z);
    *t_near = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    *t_far = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
}

TRT_DEV float4 ld4(const float4* p) { return __ldg(p); }

TRT_DEV void cswap_desc(float& ta, int& ra, float& tb, int& rb) {  // larger t first
    if (ta < tb) {
        const float t = ta; ta = tb; tb = t;
        const int r = ra; ra = rb; rb = r;
    }
}

// ---- closest hit ---------------------------------------------------------------------------
struct ClosestState {
    F3 o, d, inv;
    float d_min;
    int id;
    bool amb;
    int nsp, tsp;
    float n_t[kNodeStack];
    int n_ref[kNodeStack];
    float t_t[kTriStack];
    int t_ref[kTriStack];  // (first_tri << 2) | (count - 1)
};

TRT_DEV void closest_begin(ClosestState& s, const Ray& r) {
    s.o = r.o;
    s.d = r.d;
    s.inv = f3(ref_safe_inv(r.d.x), ref_safe_inv(r.d.y), ref_safe_inv(r.d.z));
    s.d_min = 1e20f;
    s.id = -1;
    s.amb = false;
    s.nsp = 1;
    s.tsp = 0;
    s.n_t[0] = 0.f;
    s.n_ref[0] = 0;  // root
}

TRT_DEV bool closest_done(const ClosestState& s) { return s.nsp == 0 && s.tsp == 0; }

// One round for one lane.  Safe to call when done (does nothing).
template <bool COUNT>
TRT_DEV void closest_round(const SceneDev& sc, ClosestState& s, WideCounts* wc) {
    // ---- node step -----------------------------------------------------------------------------
    int cur = kWideEmptyRef;
    {
        const float limit = s.d_min * kCullSlack;
        // keep room for up to four new triangle entries
        while (s.nsp > 0 && s.tsp <= kTriStack - 4) {
            --s.nsp;
            if (s.n_t[s.nsp] < limit) { cur = s.n_ref[s.nsp]; break; }
        }
    }
    if (cur != kWideEmptyRef) {
        const float4* np = sc.wide_nodes + (size_t)cur * 8;
        const float4 lox = ld4(np), hix = ld4(np + 1), loy = ld4(np + 2), hiy = ld4(np + 3), loz = ld4(np + 4),
                     hiz = ld4(np + 5);
        const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
        if (COUNT) wc->nodes++;
        const float limit = s.d_min * kCullSlack;
        float t0, t1, t2, t3, tf;
        int r0 = ch.x, r1 = ch.y, r2 = ch.z, r3 = ch.w;
        wide_slab(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, s.o, s.inv, &t0, &tf);
        if (!(tf >= t0 && tf > 0.f && t0 < limit) || r0 == kWideEmptyRef) t0 = -1.f;
        wide_slab(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, s.o, s.inv, &t1, &tf);
        if (!(tf >= t1 && tf > 0.f && t1 < limit) || r1 == kWideEmptyRef) t1 = -1.f;
        wide_slab(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, s.o, s.inv, &t2, &tf);
        if (!(tf >= t2 && tf > 0.f && t2 < limit) || r2 == kWideEmptyRef) t2 = -1.f;
        wide_slab(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, s.o, s.inv, &t3, &tf);
        if (!(tf >= t3 && tf > 0.f && t3 < limit) || r3 == kWideEmptyRef) t3 = -1.f;
        // A box the ray starts inside has a negative entry distance; clamp so that "missed"
        // (-1) sorts below every hit.
        t0 = t0 == -1.f ? -1.f : fmaxf(t0, 0.f);
        t1 = t1 == -1.f ? -1.f : fmaxf(t1, 0.f);
        t2 = t2 == -1.f ? -1.f : fmaxf(t2, 0.f);
        t3 = t3 == -1.f ? -1.f : fmaxf(t3, 0.f);
        // sort far -> near (misses last), then push in that order: nearest ends on top
        cswap_desc(t0, r0, t1, r1);
        cswap_desc(t2, r2, t3, r3);
        cswap_desc(t0, r0, t2, r2);
        cswap_desc(t1, r1, t3, r3);
        cswap_desc(t1, r1, t2, r2);
        const float ts[4] = {t0, t1, t2, t3};
        const int rs[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (ts[k] >= 0.f) {
                if (rs[k] >= 0) { s.n_t[s.nsp] = ts[k]; s.n_ref[s.nsp] = rs[k]; s.nsp++; }
                else { s.t_t[s.tsp] = ts[k]; s.t_ref[s.tsp] = ~rs[k]; s.tsp++; }
            }
        }
    }
    // ---- triangle step ------------------------------------------------------------------------
    int tri = -1;
    {
        const float limit = s.d_min * kCullSlack;
        while (s.tsp > 0) {
            const int top = s.tsp - 1;
            if (!(s.t_t[top] < limit)) { s.tsp = top; continue; }
            const int code = s.t_ref[top];
            tri = code >> 2;
            if (code & 3) s.t_ref[top] = code + 3;  // first + 1, count - 1
            else s.tsp = top;
            break;
        }
    }
    if (tri >= 0) {
        const float4* tp = sc.tris + (size_t)tri * 3;
        const float4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
        if (COUNT) wc->tris++;
        const float t = ref_tri_edges(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), s.o, s.d);
        const int tid = f2i(a.w);
        if (t > 0.f && (t < s.d_min || (t == s.d_min && tid < s.id))) {
            // would the reference traversal have reached this triangle?
            const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
            float entry;
            if (ref_slab(bmin, bmax, s.o, s.inv, 0.f, 3e38f, &entry)) {
                if (entry < t) { s.d_min = t; s.id = tid; }
                else s.amb = true;  // reach depends on the reference's visit order: replay
            }
        }
    }
}

// per-thread driver (test entry points; the render path uses the persistent kernels)
template <bool COUNT>
TRT_DEV int wide_closest(const SceneDev& sc, const Ray& r, float* t_out, bool* ambiguous, WideCounts* wc) {
    ClosestState s;
    closest_begin(s, r);
    while (!closest_done(s)) closest_round<COUNT>(sc, s, wc);
    *t_out = s.d_min;
    *ambiguous = s.amb;
    return s.id;
}

// ---- any hit -----------------------------------------------------------------------------------
struct ShadowState {
    F3 o, d, inv;
    float max_dist, t_hi;
    bool occluded;
    int nsp, tsp;
    int n_ref[kNodeStack];
    int t_ref[kTriStack];
};

TRT_DEV void shadow_begin(ShadowState& s, const Ray& r, float max_dist) {
    s.o = r.o;
    s.d = r.d;
    s.inv = f3(p_rcp(r.d.x), p_rcp(r.d.y), p_rcp(r.d.z));  // raw reciprocal, reference :276
    s.max_dist = max_dist;
    s.t_hi = p_sub(max_dist, 0.001f);
    s.occluded = false;
    s.nsp = 1;
    s.tsp = 0;
    s.n_ref[0] = 0;
}

TRT_DEV bool shadow_done(const ShadowState& s) { return s.occluded || (s.nsp == 0 && s.tsp == 0); }

template <bool COUNT>
TRT_DEV void shadow_round(const SceneDev& sc, ShadowState& s, WideCounts* wc) {
    if (s.occluded) return;
    if (s.nsp > 0 && s.tsp <= kTriStack - 4) {
        const int cur = s.n_ref[--s.nsp];
        const float4* np = sc.wide_nodes + (size_t)cur * 8;
        const float4 lox = ld4(np), hix = ld4(np + 1), loy = ld4(np + 2), hiy = ld4(np + 3), loz = ld4(np + 4),
                     hiz = ld4(np + 5);
        const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
        if (COUNT) wc->nodes++;
        const int rs[4] = {ch.x, ch.y, ch.z, ch.w};
        float tn[4], tf[4];
        wide_slab(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, s.o, s.inv, &tn[0], &tf[0]);
        wide_slab(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, s.o, s.inv, &tn[1], &tf[1]);
        wide_slab(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, s.o, s.inv, &tn[2], &tf[2]);
        wide_slab(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, s.o, s.inv, &tn[3], &tf[3]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (tf[k] >= tn[k] && tf[k] > 0.001f && tn[k] < s.max_dist && rs[k] != kWideEmptyRef) {
                if (rs[k] >= 0) s.n_ref[s.nsp++] = rs[k];
                else s.t_ref[s.tsp++] = ~rs[k];
            }
        }
    }
    if (s.tsp > 0) {
        const int top = s.tsp - 1;
        const int code = s.t_ref[top];
        const int tri = code >> 2;
        if (code & 3) s.t_ref[top] = code + 3;
        else s.tsp = top;
        const float4* tp = sc.tris + (size_t)tri * 3;
        const float4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
        if (COUNT) wc->tris++;
        const float t = ref_tri_edges(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), s.o, s.d);
        if (t > 0.001f && t < s.t_hi) {
            const int tid = f2i(a.w);
            const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
            if (ref_slab(bmin, bmax, s.o, s.inv, 0.001f, s.max_dist)) s.occluded = true;
        }
    }
}

template <bool COUNT>
TRT_DEV bool wide_shadow(const SceneDev& sc, const Ray& r, float max_dist, WideCounts* wc) {
    ShadowState s;
    shadow_begin(s, r, max_dist);
    while (!shadow_done(s)) shadow_round<COUNT>(sc, s, wc);
    return s.occluded;
}

}  // namespace trt
