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
#pragma once
#include "common.cuh"
#include "traverse_ref.cuh"

namespace trt {

struct WideCounts {
    uint32_t nodes, tris;
};

constexpr int kWideEmptyRef = 0x7fffffff;
constexpr int kWideStack = 64;

// reference-formula slab for one child of a wide node; near/far via fminf/fmaxf (no NaN
// can occur with a finite inverse direction; with an infinite one see DESIGN.md)
TRT_DEV void wide_slab(float lox, float hix, float loy, float hiy, float loz, float hiz, const F3 o, const F3 inv,
                       float* t_near, float* t_far) {
    const float x1 = p_mul(p_sub(lox, o.x), inv.x), x2 = p_mul(p_sub(hix, o.x), inv.x);
    const float y1 = p_mul(p_sub(loy, o.y), inv.y), y2 = p_mul(p_sub(hiy, o.y), inv.y);
    const float z1 = p_mul(p_sub(loz, o.z), inv.z), z2 = p_mul(p_sub(hiz, o.z), inv.z);
    *t_near = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    *t_far = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
}

TRT_DEV float4 ld4(const float4* p) { return __ldg(p); }

// Closest hit.  *ambiguous is set when the caller must replay the ray in reference order.
template <bool COUNT>
TRT_DEV int wide_closest(const SceneDev& sc, const Ray& r, float* t_out, bool* ambiguous, WideCounts* wc) {
    const F3 inv = f3(ref_safe_inv(r.d.x), ref_safe_inv(r.d.y), ref_safe_inv(r.d.z));
    float d_min = 1e20f;
    int id = -1;
    bool amb = false;

    float s_t[kWideStack];
    int s_ref[kWideStack];
    int sp = 0;
    int cur = 0;  // root
    for (;;) {
        if (cur >= 0) {
            // ---- inner node: test the four child boxes --------------------------------------
            const float4* np = sc.wide_nodes + (size_t)cur * 8;
            const float4 lox = ld4(np), hix = ld4(np + 1), loy = ld4(np + 2), hiy = ld4(np + 3), loz = ld4(np + 4),
                         hiz = ld4(np + 5);
            const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
            if (COUNT) wc->nodes++;
            // culling limit with slack: boxes that start marginally behind the current hit are
            // still opened so that near-ties are seen (and flagged) rather than silently skipped
            const float limit = d_min * 1.0005f;
            float tn[4], tf;
            bool hit[4];
            wide_slab(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, r.o, inv, &tn[0], &tf);
            hit[0] = tf >= tn[0] && tf > 0.f && tn[0] < limit && ch.x != kWideEmptyRef;
            wide_slab(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, r.o, inv, &tn[1], &tf);
            hit[1] = tf >= tn[1] && tf > 0.f && tn[1] < limit && ch.y != kWideEmptyRef;
            wide_slab(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, r.o, inv, &tn[2], &tf);
            hit[2] = tf >= tn[2] && tf > 0.f && tn[2] < limit && ch.z != kWideEmptyRef;
            wide_slab(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, r.o, inv, &tn[3], &tf);
            hit[3] = tf >= tn[3] && tf > 0.f && tn[3] < limit && ch.w != kWideEmptyRef;
            const int cref[4] = {ch.x, ch.y, ch.z, ch.w};
            // nearest child continues, the others go on the stack
            int best = -1;
            float best_t = 3e38f;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (hit[k] && tn[k] < best_t) { best_t = tn[k]; best = k; }
            if (best < 0) {
                cur = kWideEmptyRef;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (hit[k] && k != best) { s_t[sp] = tn[k]; s_ref[sp] = cref[k]; sp++; }
                cur = cref[best];
            }
        } else {
            // ---- leaf: exact triangle tests ------------------------------------------------------
            const int code = ~cur;
            const int first = code >> 2, count = (code & 3) + 1;
            for (int k = 0; k < count; k++) {
                const float4* tp = sc.tris + (size_t)(first + k) * 3;
                const float4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
                if (COUNT) wc->tris++;
                const float t = ref_tri_edges(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), r.o, r.d);
                if (!(t > 0.f)) continue;
                const int tid = f2i(a.w);
                if (t < d_min || (t == d_min && tid < id)) {
                    // would the reference traversal have reached this triangle?
                    const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
                    float entry;
                    if (!ref_slab(bmin, bmax, r.o, inv, 0.f, 3e38f, &entry)) continue;  // never reached
                    if (entry < t) { d_min = t; id = tid; }
                    else amb = true;  // reach depends on visit order: replay
                } else if (t < d_min * 1.0005f) {
                    // a near-tie behind the current hit: harmless unless the current hit is itself
                    // replaced later; nothing to do (kept for clarity)
                }
            }
            cur = kWideEmptyRef;
        }
        // ---- pop ------------------------------------------------------------------------------------
        while (cur == kWideEmptyRef) {
            if (sp == 0) {
                *t_out = d_min;
                *ambiguous = amb;
                return id;
            }
            sp--;
            if (s_t[sp] < d_min * 1.0005f) cur = s_ref[sp];
        }
    }
}

// Any hit in (0.001, max_dist - 0.001) with the reference's raw reciprocal direction.
template <bool COUNT>
TRT_DEV bool wide_shadow(const SceneDev& sc, const Ray& r, float max_dist, WideCounts* wc) {
    const F3 inv = f3(p_rcp(r.d.x), p_rcp(r.d.y), p_rcp(r.d.z));
    const float t_hi = p_sub(max_dist, 0.001f);
    int s_ref[kWideStack];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const float4* np = sc.wide_nodes + (size_t)cur * 8;
            const float4 lox = ld4(np), hix = ld4(np + 1), loy = ld4(np + 2), hiy = ld4(np + 3), loz = ld4(np + 4),
                         hiz = ld4(np + 5);
            const int4 ch = __ldg(reinterpret_cast<const int4*>(np + 6));
            if (COUNT) wc->nodes++;
            float tn, tf;
            int next = kWideEmptyRef;
            wide_slab(lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, r.o, inv, &tn, &tf);
            if (tf >= tn && tf > 0.001f && tn < max_dist && ch.x != kWideEmptyRef) next = ch.x;
            wide_slab(lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, r.o, inv, &tn, &tf);
            if (tf >= tn && tf > 0.001f && tn < max_dist && ch.y != kWideEmptyRef) {
                if (next != kWideEmptyRef) s_ref[sp++] = next;
                next = ch.y;
            }
            wide_slab(lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, r.o, inv, &tn, &tf);
            if (tf >= tn && tf > 0.001f && tn < max_dist && ch.z != kWideEmptyRef) {
                if (next != kWideEmptyRef) s_ref[sp++] = next;
                next = ch.z;
            }
            wide_slab(lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, r.o, inv, &tn, &tf);
            if (tf >= tn && tf > 0.001f && tn < max_dist && ch.w != kWideEmptyRef) {
                if (next != kWideEmptyRef) s_ref[sp++] = next;
                next = ch.w;
            }
            cur = next;
        } else {
            const int code = ~cur;
            const int first = code >> 2, count = (code & 3) + 1;
            for (int k = 0; k < count; k++) {
                const float4* tp = sc.tris + (size_t)(first + k) * 3;
                const float4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
                if (COUNT) wc->tris++;
                const float t = ref_tri_edges(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), r.o, r.d);
                if (t > 0.001f && t < t_hi) {
                    const int tid = f2i(a.w);
                    const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
                    if (ref_slab(bmin, bmax, r.o, inv, 0.001f, max_dist)) return true;
                }
            }
            cur = kWideEmptyRef;
        }
        if (cur == kWideEmptyRef) {
            if (sp == 0) return false;
            cur = s_ref[--sp];
        }
    }
}

}  // namespace trt
