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

TRT_DEV float4 ld4(const float4* p) { return __ldg(p); }

// Slab interval of one child from its near/far planes.  The planes are already selected by
// the sign of the inverse direction, so (near - o) * inv <= (far - o) * inv holds by the
// monotonicity of the two roundings and equals the reference's min/max of the two products.
// The interval is clamped to [0, limit]; comparing with <= instead of the reference's strict
// tests only ever adds candidates (superset).
TRT_DEV bool child_interval(float nx, float fx, float ny, float fy, float nz, float fz, const F3 o, const F3 inv,
                            float lo_clamp, float hi_clamp, float* t_near) {
    const float ax = p_mul(p_sub(nx, o.x), inv.x), bx = p_mul(p_sub(fx, o.x), inv.x);
    const float ay = p_mul(p_sub(ny, o.y), inv.y), by = p_mul(p_sub(fy, o.y), inv.y);
    const float az = p_mul(p_sub(nz, o.z), inv.z), bz = p_mul(p_sub(fz, o.z), inv.z);
    const float tn = fmaxf(fmaxf(ax, ay), fmaxf(az, lo_clamp));
    const float tf = fminf(fminf(bx, by), fminf(bz, hi_clamp));
    *t_near = tn;
    return tn <= tf;
}

// The reference's leaf-box test for a closest-hit ray, from sign-selected planes: with a finite
// inverse direction no NaN can occur, and min/max of the two plane products are then simply
// the near/far products.  Returns whether the box passes independently of d_min (exit >= entry
// and exit > 0) and the entry distance.
TRT_DEV bool leaf_box_reach(const float4 bmin, const float4 bmax, const F3 o, const F3 inv, int sx, int sy, int sz,
                            float* entry) {
    const float ax = p_mul(p_sub(sx ? bmax.x : bmin.x, o.x), inv.x), bx = p_mul(p_sub(sx ? bmin.x : bmax.x, o.x), inv.x);
    const float ay = p_mul(p_sub(sy ? bmax.y : bmin.y, o.y), inv.y), by = p_mul(p_sub(sy ? bmin.y : bmax.y, o.y), inv.y);
    const float az = p_mul(p_sub(sz ? bmax.z : bmin.z, o.z), inv.z), bz = p_mul(p_sub(sz ? bmin.z : bmax.z, o.z), inv.z);
    const float tn = fmaxf(fmaxf(ax, ay), az);
    const float tf = fminf(fminf(bx, by), bz);
    *entry = tn;
    return tf >= tn && tf > 0.f;
}

// ---- closest hit ---------------------------------------------------------------------------
// Per-lane stacks share one local array of 64-bit entries (distance bits, reference): the
// node stack grows up from 0, the triangle stack grows down from the end.
constexpr int kStackEntries = kNodeStack + kTriStack;

struct ClosestState {
    F3 o, d, inv;
    float d_min;
    int id;
    int cur;         // next inner node to open (held in a register), or kWideEmptyRef
    int sx, sy, sz;  // 1 when the inverse direction is negative on that axis
    bool amb;
    int nsp, tsp;
};
// The stack array is a separate object on purpose: a struct that contains a dynamically indexed
// array is placed in local memory as a whole, and every scalar of the ray state would then be
// loaded and stored around each round.
struct ClosestStack {
    uint2 e[kStackEntries];
};

TRT_DEV void closest_begin(ClosestState& s, const Ray& r) {
    s.o = r.o;
    s.d = r.d;
    s.inv = f3(ref_safe_inv(r.d.x), ref_safe_inv(r.d.y), ref_safe_inv(r.d.z));
    s.sx = s.inv.x < 0.f;
    s.sy = s.inv.y < 0.f;
    s.sz = s.inv.z < 0.f;
    s.d_min = 1e20f;
    s.id = -1;
    s.amb = false;
    s.nsp = 0;
    s.tsp = 0;
    s.cur = 0;  // root
}

TRT_DEV bool closest_done(const ClosestState& s) { return s.cur == kWideEmptyRef && s.nsp == 0 && s.tsp == 0; }

// One round for one lane: at most one wide-node step and one triangle step.  Written with
// selects and unconditional stores instead of branches so the warp executes one instruction
// stream for all lanes that have work of a kind.
template <bool COUNT>
TRT_DEV void closest_round(const SceneDev& sc, ClosestState& s, ClosestStack& stack, WideCounts* wc) {
    uint2* const stk = stack.e;
    const float limit = s.d_min * kCullSlack;
    // ---- choose this round's work and issue its loads --------------------------------------
    int cur = s.cur;
    if (s.tsp > kTriStack - 4) cur = kWideEmptyRef;  // keep room for four new triangle entries
    else if (cur == kWideEmptyRef) {
        while (s.nsp > 0) {
            const uint2 e = stk[--s.nsp];
            if (__uint_as_float(e.x) < limit) { cur = (int)e.y; break; }
        }
    }
    const bool do_node = cur != kWideEmptyRef;
    if (do_node) s.cur = kWideEmptyRef;
    int tri = -1;
    while (s.tsp > 0) {
        const int top = kStackEntries - s.tsp;
        const uint2 e = stk[top];
        if (!(__uint_as_float(e.x) < limit)) { s.tsp--; continue; }
        const int code = (int)e.y;
        tri = code >> 2;
        if (code & 3) stk[top].y = (unsigned)(code + 3);  // first + 1, count - 1
        else s.tsp--;
        break;
    }
    float4 nx, fx, ny, fy, nz, fz, ta, tb, tc;
    int4 ch;
    if (do_node) {
        const float4* np = sc.wide_nodes + (size_t)cur * 8;
        nx = ld4(np + s.sx);     fx = ld4(np + (s.sx ^ 1));
        ny = ld4(np + 2 + s.sy); fy = ld4(np + 2 + (s.sy ^ 1));
        nz = ld4(np + 4 + s.sz); fz = ld4(np + 4 + (s.sz ^ 1));
        ch = __ldg(reinterpret_cast<const int4*>(np + 6));
    }
#ifdef TRT_EARLY_TRI
    if (tri >= 0) {
        const float4* tp = sc.tris + (size_t)tri * 3;
        ta = ld4(tp); tb = ld4(tp + 1); tc = ld4(tp + 2);
    }
#endif
    // ---- node step ----------------------------------------------------------------------------
    if (do_node) {
        if (COUNT) wc->nodes++;
        float t[4];
        bool h[4];
        h[0] = child_interval(nx.x, fx.x, ny.x, fy.x, nz.x, fz.x, s.o, s.inv, 0.f, limit, &t[0]);
        h[1] = child_interval(nx.y, fx.y, ny.y, fy.y, nz.y, fz.y, s.o, s.inv, 0.f, limit, &t[1]);
        h[2] = child_interval(nx.z, fx.z, ny.z, fy.z, nz.z, fz.z, s.o, s.inv, 0.f, limit, &t[2]);
        h[3] = child_interval(nx.w, fx.w, ny.w, fy.w, nz.w, fz.w, s.o, s.inv, 0.f, limit, &t[3]);
        const int r[4] = {ch.x, ch.y, ch.z, ch.w};
        // nearest inner child (tournament on distances; a miss or a leaf counts as +inf)
        float ti[4];
#pragma unroll
        for (int k = 0; k < 4; k++) ti[k] = (h[k] && r[k] >= 0) ? t[k] : 3e38f;
        const bool a01 = ti[1] < ti[0], a23 = ti[3] < ti[2];
        const float t01 = a01 ? ti[1] : ti[0], t23 = a23 ? ti[3] : ti[2];
        const int k01 = a01 ? 1 : 0, k23 = a23 ? 3 : 2;
        const bool ab = t23 < t01;
        const int kbest = ab ? k23 : k01;
        const float tbest = ab ? t23 : t01;
        const int r01 = a01 ? r[1] : r[0], r23 = a23 ? r[3] : r[2];
        s.cur = tbest < 3e38f ? (ab ? r23 : r01) : kWideEmptyRef;
        int nsp = s.nsp, tsp = s.tsp;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool leaf = r[k] < 0;
            const bool push_tri = h[k] && leaf;
            const bool push_node = h[k] && !leaf && k != kbest;
            const int idx = push_tri ? (kStackEntries - 1 - tsp) : nsp;  // neither: a free slot, harmless
            stk[idx] = make_uint2(__float_as_uint(t[k]), (unsigned)(leaf ? ~r[k] : r[k]));
            nsp += push_node;
            tsp += push_tri;
        }
        s.nsp = nsp;
        s.tsp = tsp;
    }
    // ---- triangle step ------------------------------------------------------------------------
    if (tri >= 0) {
        if (COUNT) wc->tris++;
#ifndef TRT_EARLY_TRI
        const float4* tp = sc.tris + (size_t)tri * 3;
        ta = ld4(tp); tb = ld4(tp + 1); tc = ld4(tp + 2);
#endif
        const float t = ref_tri_edges(f3(ta.x, ta.y, ta.z), f3(tb.x, tb.y, tb.z), f3(tc.x, tc.y, tc.z), s.o, s.d);
        const int tid = f2i(ta.w);
        if (t > 0.f && (t < s.d_min || (t == s.d_min && tid < s.id))) {
            // would the reference traversal have reached this triangle?
            const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
            float entry;
            if (leaf_box_reach(bmin, bmax, s.o, s.inv, s.sx, s.sy, s.sz, &entry)) {
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
    ClosestStack stack;
    closest_begin(s, r);
    while (!closest_done(s)) closest_round<COUNT>(sc, s, stack, wc);
    *t_out = s.d_min;
    *ambiguous = s.amb;
    return s.id;
}

// ---- any hit -----------------------------------------------------------------------------------
struct ShadowState {
    F3 o, d, inv;
    float max_dist, t_hi;
    int sx, sy, sz;
    bool occluded;
    int nsp, tsp;
};
struct ShadowStack {
    int n_ref[kNodeStack];
    int t_ref[kTriStack];
};

TRT_DEV void shadow_begin(ShadowState& s, ShadowStack& stack, const Ray& r, float max_dist) {
    s.o = r.o;
    s.d = r.d;
    s.inv = f3(p_rcp(r.d.x), p_rcp(r.d.y), p_rcp(r.d.z));  // raw reciprocal, reference :276
    s.sx = s.inv.x < 0.f;
    s.sy = s.inv.y < 0.f;
    s.sz = s.inv.z < 0.f;
    s.max_dist = max_dist;
    s.t_hi = p_sub(max_dist, 0.001f);
    s.occluded = false;
    s.nsp = 1;
    s.tsp = 0;
    stack.n_ref[0] = 0;
}

TRT_DEV bool shadow_done(const ShadowState& s) { return s.occluded || (s.nsp == 0 && s.tsp == 0); }

template <bool COUNT>
TRT_DEV void shadow_round(const SceneDev& sc, ShadowState& s, ShadowStack& stack, WideCounts* wc) {
    if (s.occluded) return;
    int cur = kWideEmptyRef;
    if (s.nsp > 0 && s.tsp <= kTriStack - 4) cur = stack.n_ref[--s.nsp];
    int tri = -1;
    if (s.tsp > 0) {
        const int top = s.tsp - 1;
        const int code = stack.t_ref[top];
        tri = code >> 2;
        if (code & 3) stack.t_ref[top] = code + 3;
        else s.tsp = top;
    }
    float4 nx, fx, ny, fy, nz, fz, ta, tb, tc;
    int4 ch;
    if (cur != kWideEmptyRef) {
        const float4* np = sc.wide_nodes + (size_t)cur * 8;
        nx = ld4(np + s.sx);     fx = ld4(np + (s.sx ^ 1));
        ny = ld4(np + 2 + s.sy); fy = ld4(np + 2 + (s.sy ^ 1));
        nz = ld4(np + 4 + s.sz); fz = ld4(np + 4 + (s.sz ^ 1));
        ch = __ldg(reinterpret_cast<const int4*>(np + 6));
    }
#ifdef TRT_EARLY_TRI
    if (tri >= 0) {
        const float4* tp = sc.tris + (size_t)tri * 3;
        ta = ld4(tp); tb = ld4(tp + 1); tc = ld4(tp + 2);
    }
#endif
    if (cur != kWideEmptyRef) {
        if (COUNT) wc->nodes++;
        float t;
        bool h[4];
        // the reference's box interval for shadow rays is (0.001, max_dist)
        h[0] = child_interval(nx.x, fx.x, ny.x, fy.x, nz.x, fz.x, s.o, s.inv, 0.001f, s.max_dist, &t);
        h[1] = child_interval(nx.y, fx.y, ny.y, fy.y, nz.y, fz.y, s.o, s.inv, 0.001f, s.max_dist, &t);
        h[2] = child_interval(nx.z, fx.z, ny.z, fy.z, nz.z, fz.z, s.o, s.inv, 0.001f, s.max_dist, &t);
        h[3] = child_interval(nx.w, fx.w, ny.w, fy.w, nz.w, fz.w, s.o, s.inv, 0.001f, s.max_dist, &t);
        const int r[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (h[k]) {
                if (r[k] >= 0) stack.n_ref[s.nsp++] = r[k];
                else stack.t_ref[s.tsp++] = ~r[k];
            }
        }
    }
    if (tri >= 0) {
        if (COUNT) wc->tris++;
#ifndef TRT_EARLY_TRI
        const float4* tp = sc.tris + (size_t)tri * 3;
        ta = ld4(tp); tb = ld4(tp + 1); tc = ld4(tp + 2);
#endif
        const float t = ref_tri_edges(f3(ta.x, ta.y, ta.z), f3(tb.x, tb.y, tb.z), f3(tc.x, tc.y, tc.z), s.o, s.d);
        if (t > 0.001f && t < s.t_hi) {
            const int tid = f2i(ta.w);
            const float4 bmin = ld4(sc.leaf_box + (size_t)tid * 2), bmax = ld4(sc.leaf_box + (size_t)tid * 2 + 1);
            if (ref_slab(bmin, bmax, s.o, s.inv, 0.001f, s.max_dist)) s.occluded = true;
        }
    }
}

template <bool COUNT>
TRT_DEV bool wide_shadow(const SceneDev& sc, const Ray& r, float max_dist, WideCounts* wc) {
    ShadowState s;
    ShadowStack stack;
    shadow_begin(s, stack, r, max_dist);
    while (!shadow_done(s)) shadow_round<COUNT>(sc, s, stack, wc);
    return s.occluded;
}

}  // namespace trt
