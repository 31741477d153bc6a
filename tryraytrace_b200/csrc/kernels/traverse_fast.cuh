// traverse_fast.cuh -- TRAVERSE_FAST: ordered traversal of the 4-wide BVH with an exact accept
// rule, so that the result equals the reference traversal's (reference src/renderer.cu:371-425
// closest hit, :273-314 any hit) without visiting its ~60-80 nodes per ray.
//
// Why the result is the same (DESIGN.md "exactness argument"):
//  1. Superset.  Every wide box contains the reference leaf boxes below it, and the slab test
//     here is the reference's own formula -- (plane - o) * inv, a rounded subtract then a rounded
//     multiply, with the same o and inv.  Both roundings are monotone, so a wider box can only
//     give an earlier entry and a later exit: whenever the reference's test of a leaf box passes,
//     the tests of all wide boxes above that triangle pass too.
//  2. Same triangle arithmetic.  Candidates are tested with the reference's Moeller-Trumbore
//     operation sequence on the original vertices (edges are re-subtracted per test, exactly as
//     reference :239-240 does).
//  3. Exact reachability.  The reference tests triangle k only if the slab tests of all its
//     ancestors pass.  Ancestor boxes contain the leaf box, so by (1) they pass whenever the leaf
//     box passes with entry < t_k; a candidate whose reference leaf box fails outright can never
//     be reached by the reference and is dropped.  The leaf box is re-derived from the three
//     vertices with the builder's rule (reference src/bvh.cpp:12-30); the upload verifies per
//     triangle that this reproduces the uploaded leaf node bit for bit and flags the triangle
//     otherwise (kTriNoDerive), which forces the replay below.
//  4. Ambiguity.  The only order-dependent case left is a candidate whose leaf-box entry is not
//     below its own hit distance (the reference culls against the running d_min, which depends
//     on visit order).  Such rays are re-run through TRAVERSE_REF.  Ties in t resolve to the
//     lowest object index, as the reference's ascending leaf order does.
//  5. Deferred verification.  (3) and (4) only matter for the FINAL winner w = the candidate with
//     the smallest t (lowest id on ties) over the superset: the reference's running d_min never
//     drops below t_w (every triangle it can accept is a candidate here), so if w's leaf box
//     passes with entry < t_w all of w's ancestors pass whenever the reference visits them, w is
//     tested, and nothing tested beats it.  So candidates are accepted on the triangle test
//     alone while traversing, and the leaf-box reach of the winner is checked once per ray; if
//     it fails (FP corner cases, or a box that cannot be re-derived) the ray is replayed.  The
//     check runs where the hit is CONSUMED (resolve_hit below: the shade kernel, or the unpack
//     kernel of the parity entry points), one thread per ray at full lane occupancy, not inside
//     the traversal loop where only the few lanes that just finished would take part.
// Any-hit queries have a fixed interval, so (3) decides them exactly and no replay exists.
//
// Execution model (kernels in wavefront.cu).  One persistent CTA per SM.  Two levels:
//   * TOP PHASE.  A warp takes a chunk of 32 consecutive pool slots, one ray per lane, and tests
//     the root-level list (the oversized primitives and the light sources the builder lifted out
//     of the tree, <= 12, in the constant bank) brute force and fully converged, plus the tree's
//     bounding box.  In the benchmark scenes 60-80% of all rays are decided right there (they only
//     ever see the room's walls) at 100% lane utilisation and without touching a stack.
//   * TREE PHASE.  Rays that enter the tree are compacted (__ballot_sync/__popc) into a per-warp
//     queue in shared memory, carrying the d_min / id found so far; idle traversal lanes refill
//     from that queue.
// The tree traversal is speculative while-while (Aila & Laine 2009) in warp-converged PHASES:
// a few node steps (four child-box tests each, packed FADD2/FMUL2), then triangle steps for the
// leaves that piled up meanwhile, so triangle tests run with fuller warps.  Inner
// children go to a per-lane node stack, leaf children to a separate per-lane triangle stack;
// both live in SHARED memory (lane-interleaved, conflict free), with a local-memory overflow
// that only deep trees ever touch.  The top of the tree (the first k nodes, emitted by the
// builder in descending box area) is staged into shared memory once per CTA.  Rays arrive in
// chunks of 32 through a per-warp double buffer filled by TMA bulk copies (cp.async.bulk +
// mbarrier), so the refill of idle lanes never waits on HBM.
#pragma once
#include "common.cuh"
#include "traverse_ref.cuh"

namespace trt {

struct WideCounts {
    uint32_t nodes, tris;
};

constexpr int kWideEmptyRef = 0x7fffffff;
constexpr int kSmemNodeStride = 112;  // bytes between staged nodes in shared memory: the 7 used float4 of a node
                                      // (the pad word is not staged); 112 = 28 banks, so consecutive nodes start in
                                      // 8 different 16-byte bank groups
constexpr int kSpillEntries = 128;  // logical node stack bound: 3 * wide depth + 1 (checked at upload)
constexpr float kCullSlack = 1.0005f;
constexpr int kTriIdMask = 0x3fffffff;
constexpr int kTriNoDerive = 0x40000000;  // leaf box cannot be re-derived from the vertices: replay

// ---- async staging primitives (sm_90+/sm_100 PTX) ----------------------------------------------
TRT_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

TRT_DEV void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
TRT_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TRT_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
TRT_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
TRT_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// Per-lane traversal stacks live in shared memory and are addressed with 32-bit shared-window
// addresses.  All stack accesses are volatile asm WITHOUT a memory clobber: they stay ordered
// among themselves (pushes and pops never swap), while the compiler remains free to schedule
// the global node/triangle loads across them.  Pushes are predicated stores, not branches.
TRT_DEV void sts64_if(bool p, uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q st.shared.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((uint32_t)p),
        "r"(addr), "r"(x), "r"(y));
}
TRT_DEV void sts32_if(bool p, uint32_t addr, uint32_t x) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q st.shared.u32 [%1], %2;\n\t}" ::"r"((uint32_t)p),
        "r"(addr), "r"(x));
}
TRT_DEV uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
TRT_DEV uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Both pushes of one child in a single predicated sequence: the leaf push (triangle stack,
// grows down) when `hit && ref < 0`, the inner push (node stack, grows up) when the child's key
// is valid and not the one kept in a register.  np / tp are bumped in place under the same
// predicates, so a child costs nine instructions and no branch.
template <uint32_t E>
TRT_DEV void push_child(uint32_t& np, uint32_t& tp, float tn, float tf, int ref, unsigned key, unsigned best) {
    asm volatile(
        "{\n\t.reg .pred pt, pn;\n\t.reg .b32 nref;\n\t"
        "setp.le.ftz.f32 pt, %2, %3;\n\t"
        "setp.lt.and.s32 pt, %4, 0, pt;\n\t"
        "setp.ne.u32 pn, %5, %6;\n\t"
        "setp.ne.and.u32 pn, %5, 0xffffffff, pn;\n\t"
        "not.b32 nref, %4;\n\t"
        "@pt st.shared.v2.b32 [%1], {%7, nref};\n\t"
        "@pt sub.u32 %1, %1, %8;\n\t"
        "@pn st.shared.v2.b32 [%0], {%7, %4};\n\t"
        "@pn add.u32 %0, %0, %8;\n\t}"
        : "+r"(np), "+r"(tp)
        : "f"(tn), "f"(tf), "r"(ref), "r"(key), "r"(best), "r"(__float_as_uint(tn)), "n"(E));
}
// any hit: 32-bit entries, no distances, every hit inner child is pushed
template <uint32_t E>
TRT_DEV void push_child_any(uint32_t& np, uint32_t& tp, float tn, float tf, int ref) {
    asm volatile(
        "{\n\t.reg .pred h, pt, pn;\n\t.reg .b32 nref;\n\t"
        "setp.le.ftz.f32 h, %2, %3;\n\t"
        "setp.lt.and.s32 pt, %4, 0, h;\n\t"
        "setp.ge.and.s32 pn, %4, 0, h;\n\t"
        "not.b32 nref, %4;\n\t"
        "@pt st.shared.b32 [%1], nref;\n\t"
        "@pt sub.u32 %1, %1, %5;\n\t"
        "@pn st.shared.b32 [%0], %4;\n\t"
        "@pn add.u32 %0, %0, %5;\n\t}"
        : "+r"(np), "+r"(tp)
        : "f"(tn), "f"(tf), "r"(ref), "n"(E));
}

// Packed FP32 (sm_100 FADD2 / FMUL2): two lanes of the same rn/ftz operation per instruction,
// each lane rounded exactly like the scalar op, so the slab arithmetic below is bit-identical
// to (plane - o) * inv evaluated per child.
TRT_DEV float2 sub2(float2 a, float s) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
        "sub.rn.ftz.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
TRT_DEV float2 mul2(float2 a, float s) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
        "mul.rn.ftz.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
// pair x pair / pair x scalar forms used by the two-triangle test (tri_test_pair)
TRT_DEV float2 mul2p(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.ftz.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
TRT_DEV float2 add2(float2 a, float s) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
        "add.rn.ftz.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
TRT_DEV float2 sub2p(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "sub.rn.ftz.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
TRT_DEV float2 fma2p(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.ftz.f32x2 rr, ra, rb, rc;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
TRT_DEV float2 fma2s(float2 a, float s, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.ftz.f32x2 rr, ra, rb, rc;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
    return r;
}
// (plane - o) * inv for the four children of one plane vector
TRT_DEV float4 plane_t(float4 p, float o, float inv) {
    const float2 a = mul2(sub2(make_float2(p.x, p.y), o), inv);
    const float2 b = mul2(sub2(make_float2(p.z, p.w), o), inv);
    return make_float4(a.x, a.y, b.x, b.y);
}

// streaming (evict-first) stores for results that are read once by the next kernel
TRT_DEV void st_cs_f2(float2* p, float2 v) { __stcs(p, v); }

// ---- geometry helpers --------------------------------------------------------------------------
// The reference leaf box of a triangle: bounds of the three vertices, padded by 1e-3 on axes
// thinner than 1e-3 (reference src/bvh.cpp:12-30).  min/max of finite values do not depend on
// the order or on the instruction used; the three adds are single roundings like the host's.
TRT_DEV void derive_leaf_box(const F3 v0, const F3 v1, const F3 v2, float4* bmin, float4* bmax) {
    const float pad = 1e-3f;
    float lx = fminf(fminf(v0.x, v1.x), v2.x), hx = fmaxf(fmaxf(v0.x, v1.x), v2.x);
    float ly = fminf(fminf(v0.y, v1.y), v2.y), hy = fmaxf(fmaxf(v0.y, v1.y), v2.y);
    float lz = fminf(fminf(v0.z, v1.z), v2.z), hz = fmaxf(fmaxf(v0.z, v1.z), v2.z);
    if (p_sub(hx, lx) < pad) { lx = p_sub(lx, pad); hx = p_add(hx, pad); }
    if (p_sub(hy, ly) < pad) { ly = p_sub(ly, pad); hy = p_add(hy, pad); }
    if (p_sub(hz, lz) < pad) { lz = p_sub(lz, pad); hz = p_add(hz, pad); }
    *bmin = make_float4(lx, ly, lz, 0.f);
    *bmax = make_float4(hx, hy, hz, 0.f);
}

// The reference's leaf-box test for a closest-hit ray, from sign-selected planes: with a finite
// inverse direction no NaN can occur, and min/max of the two plane products are then simply the
// near/far products.  Returns whether the box passes independently of d_min (exit >= entry and
// exit > 0) and the entry distance.
TRT_DEV bool leaf_box_reach(const float4 bmin, const float4 bmax, const F3 o, const F3 inv, bool sx, bool sy, bool sz,
                            float* entry) {
    const float ax = p_mul(p_sub(sx ? bmax.x : bmin.x, o.x), inv.x), bx = p_mul(p_sub(sx ? bmin.x : bmax.x, o.x), inv.x);
    const float ay = p_mul(p_sub(sy ? bmax.y : bmin.y, o.y), inv.y), by = p_mul(p_sub(sy ? bmin.y : bmax.y, o.y), inv.y);
    const float az = p_mul(p_sub(sz ? bmax.z : bmin.z, o.z), inv.z), bz = p_mul(p_sub(sz ? bmin.z : bmax.z, o.z), inv.z);
    const float tn = fmaxf(fmaxf(ax, ay), az);
    const float tf = fminf(fminf(bx, by), bz);
    *entry = tn;
    return tf >= tn && tf > 0.f;
}

// Moeller-Trumbore with the reference's operation sequence (traverse_ref.cuh ref_tri_edges),
// evaluated without branches: every early-out of reference :235-268 becomes one term of `ok`,
// with the same comparisons, so NaN/inf intermediates are rejected exactly where the reference
// rejects them.
TRT_DEV float tri_test_flat(const F3 v0, const F3 e1, const F3 e2, const F3 o, const F3 d) {
    const float eps = 1e-5f;
    const F3 h = x_cross(d, e2);
    const float a = x_dot(e1, h);
    const float f = p_rcp(a);
    const F3 s = x_sub(o, v0);
    const float u = p_mul(f, x_dot(s, h));
    const F3 q = x_cross(s, e1);
    const float v = p_mul(f, x_dot(d, q));
    const float t = p_mul(f, x_dot(e2, q));
    const bool ok = !(a > -eps && a < eps) && !(u < 0.f || u > 1.f) && !(v < 0.f || p_add(u, v) > 1.f) && t > eps;
    return ok ? t : 0.f;
}

// The same test for TWO triangles at once (sm_100 packed FP32: FFMA2 / FMUL2 / FADD2).  Every packed lane is
// rounded exactly like the scalar instruction, and each line below is the packed form of the corresponding line
// of tri_test_flat with the same operand roles: multiplication commutes, o + (-v0) == o - v0, and the negated
// rounded product -(x * y) of the reference's cross product is the product with one operand negated.
// nv0 = -v0, ne1 = -e1, nd = -d.  Returns t per triangle, 0 = no hit.
TRT_DEV float2 tri_test_pair(const float2 nv0[3], const float2 e1[3], const float2 ne1[3], const float2 e2[3], const F3 o,
                             const F3 d, const F3 nd) {
    const float eps = 1e-5f;
    // h = cross(d, e2)
    const float2 hx = fma2s(e2[2], d.y, mul2(e2[1], nd.z));
    const float2 hy = fma2s(e2[0], d.z, mul2(e2[2], nd.x));
    const float2 hz = fma2s(e2[1], d.x, mul2(e2[0], nd.y));
    // a = dot(e1, h)
    const float2 a = fma2p(e1[2], hz, fma2p(e1[0], hx, mul2p(e1[1], hy)));
    const float2 f = make_float2(p_rcp(a.x), p_rcp(a.y));
    // s = o - v0
    const float2 sx = add2(nv0[0], o.x), sy = add2(nv0[1], o.y), sz = add2(nv0[2], o.z);
    const float2 u = mul2p(f, fma2p(sz, hz, fma2p(sx, hx, mul2p(sy, hy))));
    // q = cross(s, e1)
    const float2 qx = fma2p(sy, e1[2], mul2p(sz, ne1[1]));
    const float2 qy = fma2p(sz, e1[0], mul2p(sx, ne1[2]));
    const float2 qz = fma2p(sx, e1[1], mul2p(sy, ne1[0]));
    const float2 v = mul2p(f, fma2s(qz, d.z, fma2s(qx, d.x, mul2(qy, d.y))));
    const float2 t = mul2p(f, fma2p(e2[2], qz, fma2p(e2[0], qx, mul2p(e2[1], qy))));
    const bool ok0 = !(a.x > -eps && a.x < eps) && !(u.x < 0.f || u.x > 1.f) && !(v.x < 0.f || p_add(u.x, v.x) > 1.f) && t.x > eps;
    const bool ok1 = !(a.y > -eps && a.y < eps) && !(u.y < 0.f || u.y > 1.f) && !(v.y < 0.f || p_add(u.y, v.y) > 1.f) && t.y > eps;
    return make_float2(ok0 ? t.x : 0.f, ok1 ? t.y : 0.f);
}

// Two leaf triangles from their records (v0|id, v1, v2): edges, negations and the test, all packed.
// e1 = v1 - v0 and e2 = v2 - v0 are the reference's single rounded subtractions (:239-240); -v0 and -e1 are
// exact sign flips (multiplication by -1).
TRT_DEV float2 tri_test_pair_records(const float4 a0, const float4 b0, const float4 c0, const float4 a1, const float4 b1,
                                     const float4 c1, const F3 o, const F3 d, const F3 nd) {
    const float2 v0[3] = {make_float2(a0.x, a1.x), make_float2(a0.y, a1.y), make_float2(a0.z, a1.z)};
    const float2 e1[3] = {sub2p(make_float2(b0.x, b1.x), v0[0]), sub2p(make_float2(b0.y, b1.y), v0[1]),
                          sub2p(make_float2(b0.z, b1.z), v0[2])};
    const float2 e2[3] = {sub2p(make_float2(c0.x, c1.x), v0[0]), sub2p(make_float2(c0.y, c1.y), v0[1]),
                          sub2p(make_float2(c0.z, c1.z), v0[2])};
    const float2 nv0[3] = {mul2(v0[0], -1.f), mul2(v0[1], -1.f), mul2(v0[2], -1.f)};
    const float2 ne1[3] = {mul2(e1[0], -1.f), mul2(e1[1], -1.f), mul2(e1[2], -1.f)};
    return tri_test_pair(nv0, e1, ne1, e2, o, d, nd);
}

// Slab interval of one child from its near/far planes (selected by the sign of the inverse
// direction when the node was loaded), clamped to [lo_clamp, hi_clamp]; `<=` instead of the
// reference's strict tests only ever adds candidates (superset).
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

// One 4-wide node: near/far plane vectors per axis and the child references.
struct NodeData {
    float4 nx, fx, ny, fy, nz, fz;
    int4 ch;
};

// nxo / nyo / nzo are the byte offsets of the near-plane vectors inside the 128-byte node:
// x: 0 or 16, y: 32 or 48, z: 64 or 80; the far plane is the other one of each pair (^16).  They are derived from
// the sign of the reciprocal direction at every node step (three selects) rather than carried in three registers:
// the traversal kernels are issue bound and want warps, and 70 registers allow 28 warps per SM where 77 allowed 24.
// The first k_smem nodes are staged in shared memory 112 bytes apart (the 16-byte plane vector a
// lane reads then falls into bank group (k - node) mod 8, so lanes reading the same k of
// different nodes spread over the banks), the rest sit in global memory 128 bytes apart.  Both
// are read through ONE generic-address code path: a warp whose lanes are split between the two
// spaces issues the seven loads once, not twice.
// Compressed node, 64 bytes (north-star subsystem 1, "compressed wide BVH"; after Ylitie, Karras & Laine 2017): the
// child boxes of a node are stored as 8-bit grid coordinates relative to the node's own box,
//     plane = fma(float(2^23 + q), scale_axis, base_axis),      q in [0, 255], scale a power of two,
// i.e. the byte is dropped into the mantissa of 2^23 (one PRMT) and one FMA decodes it; base = lo - 2^23 * scale.
// The converter (k_compress_nodes below) evaluates exactly this expression and moves q outward until the decoded
// lo plane is <= and the decoded hi plane is >= the plane of the 128-byte node, so a compressed box CONTAINS the
// uncompressed one: the superset argument at the top of this file is untouched, the ray only meets a few more
// candidates.  Layout, as four 16-byte words:
//     w0 = base.x base.y base.z scale.x | w1 = scale.y scale.z qlo_x qhi_x | w2 = qlo_y qhi_y qlo_z qhi_z | w3 = child[4]
// (q words: child k in byte k).  Two 256-bit loads per node step instead of seven 128-bit ones, and half the
// bytes: for trees far larger than the caches the traversal is bound by L1 requests (C5: 94-96 % of the L1TEX
// request rate with the 128-byte node), not by arithmetic.
// An unused child slot keeps the reference kWideEmptyRef and is made to miss explicitly (its far plane on x
// becomes -inf * sign): an inverted box cannot be relied on here, the grid has no infinity.
struct CNode {
    float base[3];
    float scale[3];
    uint32_t qx_lo, qx_hi, qy_lo, qy_hi, qz_lo, qz_hi;
    int child[4];
};
static_assert(sizeof(CNode) == 64, "CNode layout");

TRT_DEV float cnode_plane(uint32_t q_word, int k, float scale, float base) {
    // byte k of q_word into the low byte of 0x4B000000 (= 2^23 as a float): the float 2^23 + q
    const uint32_t bits = __byte_perm(q_word, 0x4B000000u, 0x7440u + (uint32_t)k);
    return p_fma(__uint_as_float(bits), scale, base);
}
TRT_DEV float4 cnode_planes(uint32_t q_word, float scale, float base) {
    return make_float4(cnode_plane(q_word, 0, scale, base), cnode_plane(q_word, 1, scale, base),
                       cnode_plane(q_word, 2, scale, base), cnode_plane(q_word, 3, scale, base));
}
// line of a compressed node into L1 ahead of the node step that will read it (no register, no wait)
TRT_DEV void prefetch_cnode(const SceneDev& sc, int node, bool p) {
    if (p) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const unsigned char*>(sc.cnodes) + (size_t)node * 64));
}
TRT_DEV void ldg256u(const unsigned char* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}

// WIDE = false: the generic path above (scenes whose tree top fits the staged / cached part).
// WIDE = true: the compressed 64-byte node, nothing staged in shared memory.  The near / far choice is made on the
// packed grid words (one select per axis), then the eight planes of an axis are decoded.
template <bool WIDE>
TRT_DEV void load_node(NodeData& n, const unsigned char* s_nodes, int k_smem, const SceneDev& sc, int node, int nxo,
                       int nyo, int nzo) {
    if (WIDE) {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(sc.cnodes) + (size_t)node * 64;
        uint4 w0, w1, w2, w3;
        ldg256u(b, w0, w1);
        ldg256u(b + 32, w2, w3);
        const float bx = __uint_as_float(w0.x), by = __uint_as_float(w0.y), bz = __uint_as_float(w0.z);
        const float sx = __uint_as_float(w0.w), sy = __uint_as_float(w1.x), sz = __uint_as_float(w1.y);
        const bool negx = nxo != 0, negy = nyo != 32, negz = nzo != 64;  // the ray travels towards -axis: near = hi
        n.nx = cnode_planes(negx ? w1.w : w1.z, sx, bx);
        n.fx = cnode_planes(negx ? w1.z : w1.w, sx, bx);
        n.ny = cnode_planes(negy ? w2.y : w2.x, sy, by);
        n.fy = cnode_planes(negy ? w2.x : w2.y, sy, by);
        n.nz = cnode_planes(negz ? w2.w : w2.z, sz, bz);
        n.fz = cnode_planes(negz ? w2.z : w2.w, sz, bz);
        n.ch = make_int4((int)w3.x, (int)w3.y, (int)w3.z, (int)w3.w);
        // unused slots: far plane = -inf * sign(inv.x), so (far - o) * inv = -inf whatever the ray
        const float kill = negx ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
        n.fx.x = n.ch.x == kWideEmptyRef ? kill : n.fx.x;
        n.fx.y = n.ch.y == kWideEmptyRef ? kill : n.fx.y;
        n.fx.z = n.ch.z == kWideEmptyRef ? kill : n.fx.z;
        n.fx.w = n.ch.w == kWideEmptyRef ? kill : n.fx.w;
        return;
    }
    const bool staged = node < k_smem;
    const unsigned char* b = staged ? s_nodes + (size_t)node * kSmemNodeStride
                                    : reinterpret_cast<const unsigned char*>(sc.wide_nodes) + (size_t)node * 128;
    n.nx = *reinterpret_cast<const float4*>(b + nxo);
    n.fx = *reinterpret_cast<const float4*>(b + (nxo ^ 16));
    n.ny = *reinterpret_cast<const float4*>(b + nyo);
    n.fy = *reinterpret_cast<const float4*>(b + (nyo ^ 16));
    n.nz = *reinterpret_cast<const float4*>(b + nzo);
    n.fz = *reinterpret_cast<const float4*>(b + (nzo ^ 16));
    n.ch = *reinterpret_cast<const int4*>(b + 96);
}

// ---- closest hit -------------------------------------------------------------------------------
// Stack geometry of one lane: S entries of E bytes, entry j at base + j * E (E = entry size x
// CTA threads, so a warp's accesses are conflict free).  Node entries grow up from `base`,
// triangle entries grow down from `base + (S-1) * E`; np / tp are the next free entries.
struct ClosestRay {
    F3 o, d, inv;
    float d_min;
    int id;
    int cur;            // next inner node to open (held in a register), or kWideEmptyRef
    uint32_t np, tp;    // shared-window addresses of the next free node / triangle entry
    int nspill;         // entries in the local overflow
};  // `id` is the winner's object index | kTriNoDerive (as stored in its record), -1 = none

// Tree-phase entry: the ray with the d_min / id / ambiguity the top phase found.
TRT_DEV void closest_begin(ClosestRay& s, const float4 o4, const float4 d4, float d_min, int id, uint32_t base,
                           uint32_t ttop) {
    s.o = f3(o4.x, o4.y, o4.z);
    s.d = f3(d4.x, d4.y, d4.z);
    s.inv = f3(ref_safe_inv(s.d.x), ref_safe_inv(s.d.y), ref_safe_inv(s.d.z));
    s.d_min = d_min;
    s.id = id;
    s.np = base;
    s.tp = ttop;
    s.nspill = 0;
    s.cur = 0;  // root
}

// TOP PHASE, closest hit: the root-level list, brute force, same triangle arithmetic and tie rule
// as the tree phase, then the tree's bounding box against [0, d_min].  Every lane of the warp
// runs this for its own ray.
struct TopResult {
    float d_min;
    int id;       // object index | kTriNoDerive, -1 none
    bool enters;  // the ray can reach something in the tree
};
TRT_DEV TopResult top_closest(const TopPrims& top, const F3 o, const F3 d) {
    const F3 inv = f3(ref_safe_inv(d.x), ref_safe_inv(d.y), ref_safe_inv(d.z));
    const bool sx = inv.x < 0.f, sy = inv.y < 0.f, sz = inv.z < 0.f;
    TopResult r;
    r.d_min = 1e20f;
    r.id = -1;
    const F3 nd = f3(-d.x, -d.y, -d.z);
    // two primitives per pass, in ascending object index: on a tie in t the earlier (lower) index stays
#pragma unroll 1
    for (int j = 0; j < top.n_pairs; j++) {
        const float2 t = tri_test_pair(top.pair_nv0[j], top.pair_e1[j], top.pair_ne1[j], top.pair_e2[j], o, d, nd);
        const int2 tid = top.pair_id[j];
        if (t.x > 0.f && t.x < r.d_min) { r.d_min = t.x; r.id = tid.x; }
        if (t.y > 0.f && t.y < r.d_min) { r.d_min = t.y; r.id = tid.y; }
    }
    float tn;
    const float limit = r.d_min * kCullSlack;
    r.enters = child_interval(sx ? top.root_hi.x : top.root_lo.x, sx ? top.root_lo.x : top.root_hi.x,
                              sy ? top.root_hi.y : top.root_lo.y, sy ? top.root_lo.y : top.root_hi.y,
                              sz ? top.root_hi.z : top.root_lo.z, sz ? top.root_lo.z : top.root_hi.z, o, inv, 0.f,
                              limit, &tn);
    return r;
}

// TOP PHASE, closest hit, RANKED.  The brute-force pass above runs the full triangle test for every
// root-level primitive (seven per ray in the benchmark scenes) although a ray inside the room can only hit
// the one or two walls in front of it.  Here every lane first ranks the primitives by where its ray enters
// their reference leaf box -- on the box's thinnest axis only (two products, exactly the reference's; flat
// wall boxes are 2e-3 thick, so this is the plane distance), or on all three axes for primitives flagged small
// (lights) -- and then tests them nearest first until no untested primitive's entry lies below d_min * slack.
// Skipping is exact for the same reasons as in the tree phase:
//   * exit <= 0 (or an empty interval): the reference's slab test of that leaf box fails (its tmax is at most
//     this axis' exit; no NaN can arise because safe_inv keeps the reciprocals finite), so the reference
//     never tests the triangle either;
//   * entry >= d_min * slack: the triangle lies inside its leaf box, its hit distance is not below the box
//     entry (up to rounding far smaller than the slack), so it cannot beat the winner.  The winner itself is
//     verified where it is consumed (resolve_hit), as always.
// The per-lane triangle operands come from a shared-memory copy of the list (s_top: v0|id, e1, e2 per
// primitive); lanes usually test one primitive (the wall their ray hits), the warp leaves the loop when no
// lane has a candidate left.
// Ranking pass: the three nearest candidates (keys k1 <= k2 <= k3, 0xffffffff = none).  FIRST: all primitives;
// otherwise only those not in `tested` whose box entry lies below `limit`.
template <bool FIRST>
TRT_DEV void top_rank(const TopPrims& top, const F3 o, const F3 inv, unsigned tested, float limit, unsigned& k1,
                      unsigned& k2, unsigned& k3) {
    k1 = k2 = k3 = 0xffffffffu;
    auto insert = [&](int p, float tn, bool ok) {
        // key: entry distance (>= 0, so the bit pattern orders like the value) with the index in the low bits
        if (!FIRST) ok = ok && tn < limit && !((tested >> p) & 1u);
        const unsigned k = ok ? ((__float_as_uint(fmaxf(tn, 0.f)) & ~15u) | (unsigned)p) : 0xffffffffu;
        const unsigned a = max(k1, k);
        k1 = min(k1, k);
        const unsigned b = max(k2, a);
        k2 = min(k2, a);
        k3 = min(k3, b);
    };
    auto thin = [&](int p, float oa, float ia) {
        const float2 t = mul2(sub2(top.thin2[p], oa), ia);  // the reference's two products on this axis
        insert(p, fminf(t.x, t.y), fmaxf(t.x, t.y) > 0.f);
    };
    int p = 0;
#pragma unroll 1
    for (const int e = top.n_thin[0]; p < e; p++) thin(p, o.x, inv.x);
    p = top.n_axis[0];
#pragma unroll 1
    for (const int e = p + top.n_thin[1]; p < e; p++) thin(p, o.y, inv.y);
    p = top.n_axis[0] + top.n_axis[1];
#pragma unroll 1
    for (const int e = p + top.n_thin[2]; p < e; p++) thin(p, o.z, inv.z);
#pragma unroll 1
    for (int i = 0; i < top.n_full; i++) {  // small primitives: the whole leaf box
        const int q = top.full_idx[i];
        const float4 lo = top.bmin[q], hi = top.bmax[q];
        const float2 x = mul2(sub2(make_float2(lo.x, hi.x), o.x), inv.x);
        const float2 y = mul2(sub2(make_float2(lo.y, hi.y), o.y), inv.y);
        const float2 z = mul2(sub2(make_float2(lo.z, hi.z), o.z), inv.z);
        const float tn = fmaxf(fmaxf(fminf(x.x, x.y), fminf(y.x, y.y)), fminf(z.x, z.y));
        const float tf = fminf(fminf(fmaxf(x.x, x.y), fmaxf(y.x, y.y)), fmaxf(z.x, z.y));
        insert(q, tn, tf > 0.f && tf >= tn);
    }
}

TRT_DEV TopResult top_closest_ranked(const TopPrims& top, const float4* s_top, const F3 o, const F3 d, bool live) {
    const F3 inv = f3(ref_safe_inv(d.x), ref_safe_inv(d.y), ref_safe_inv(d.z));
    const bool sx = inv.x < 0.f, sy = inv.y < 0.f, sz = inv.z < 0.f;
    TopResult r;
    r.d_min = 1e20f;
    r.id = -1;
    unsigned k1, k2, k3, tested = 0;
    top_rank<true>(top, o, inv, 0u, 0.f, k1, k2, k3);
    if (!live) k1 = 0xffffffffu;
    bool had3 = k3 != 0xffffffffu;  // there may be candidates beyond the three ranked ones
#pragma unroll 1
    for (;;) {
        // the nearest untested candidate, if it can still beat (or tie) the winner; the candidates are sorted,
        // so when it cannot, nothing behind it can
        const bool go = k1 != 0xffffffffu && __uint_as_float(k1 & ~15u) < r.d_min * kCullSlack;
        if (!__any_sync(0xffffffffu, go)) break;
        bool again = false;
        if (go) {
            const int p = (int)(k1 & 15u);
            const float4 a = s_top[p * 3], b = s_top[p * 3 + 1], c = s_top[p * 3 + 2];
            const float t = tri_test_flat(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), o, d);
            const int tid = f2i(a.w);
            if (t > 0.f && (t < r.d_min || (t == r.d_min && (tid & kTriIdMask) < (r.id & kTriIdMask)))) {
                r.d_min = t;
                r.id = tid;
            }
            tested |= 1u << p;
            k1 = k2;
            k2 = k3;
            k3 = 0xffffffffu;
            again = k1 == 0xffffffffu && had3;  // used up the ranked three and there may be more: rank the rest (rare)
        } else {
            k1 = 0xffffffffu;
            had3 = false;
        }
        if (__any_sync(0xffffffffu, again)) {
            unsigned n1, n2, n3;
            top_rank<false>(top, o, inv, tested, r.d_min * kCullSlack, n1, n2, n3);
            if (again) { k1 = n1; k2 = n2; k3 = n3; had3 = n3 != 0xffffffffu; }
        }
    }
    float tn;
    const float limit = r.d_min * kCullSlack;
    r.enters = child_interval(sx ? top.root_hi.x : top.root_lo.x, sx ? top.root_lo.x : top.root_hi.x,
                              sy ? top.root_hi.y : top.root_lo.y, sy ? top.root_lo.y : top.root_hi.y,
                              sz ? top.root_hi.z : top.root_lo.z, sz ? top.root_lo.z : top.root_hi.z, o, inv, 0.f,
                              limit, &tn);
    return r;
}

// Deferred verification of a ray's winner (exactness argument, point 5): its reference leaf box
// must pass the reference's slab test with an entry below the hit distance; otherwise -- or when
// the box cannot be re-derived from the vertices -- the ray is re-run in reference order.
// `id` comes in as written by the traversal kernel (object index | kTriNoDerive, < 0 = miss) and
// leaves as the final object index.  Returns true when the ray was replayed.
TRT_DEV bool resolve_hit(const SceneDev& sc, const F3 o, const F3 d, float& t, int& id) {
    if (id < 0) return false;  // a miss is a miss
    bool ok = !(id & kTriNoDerive);
    id &= kTriIdMask;
    if (ok) {
        const float4* op = sc.objects + (size_t)id * 7;
        const float4 a = __ldg(op), b = __ldg(op + 1), c = __ldg(op + 2);
        float4 bmin, bmax;
        derive_leaf_box(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), &bmin, &bmax);
        const F3 inv = f3(ref_safe_inv(d.x), ref_safe_inv(d.y), ref_safe_inv(d.z));
        float entry;
        ok = leaf_box_reach(bmin, bmax, o, inv, inv.x < 0.f, inv.y < 0.f, inv.z < 0.f, &entry) && entry < t;
    }
    if (ok || sc.n_ref_nodes == 0) return false;  // no reference tree uploaded: nothing to replay through
    Ray r;
    r.o = o;
    r.d = d;
    VisitCounts vc = {0, 0, 0};
    id = ref_closest<false>(sc, r, &t, &vc);
    return true;
}

// The traversal of one lane is cut into NODE steps and TRIANGLE steps; the kernels run them in
// phases (several node steps, then triangle steps) with all lanes of the warp taking part -- a
// lane without work of a kind falls through.  E = bytes between consecutive stack entries of
// this lane, S = entries per lane.
template <uint32_t E, int S>
TRT_DEV bool closest_has_room(const ClosestRay& s) {
    // at least four free entries (a node step can push that many); tp - np = (free - 1) * E, signed:
    // a completely full stack gives -E
    return (int)(s.tp - s.np) >= (int)(3 * E);
}
TRT_DEV bool closest_node_work(const ClosestRay& s, uint32_t base) {
    return s.cur != kWideEmptyRef || s.np != base || s.nspill > 0;
}

template <uint32_t E, int S, bool COUNT, bool WIDE>
TRT_DEV void closest_node_step(const unsigned char* s_nodes, int k_smem, const SceneDev& sc, ClosestRay& s,
                               uint32_t base, uint2* spill, WideCounts* wc) {
    const uint32_t ttop = base + (S - 1) * E;
    const float limit = s.d_min * kCullSlack;
    // ---- choose the node: the held one, else the nearest-first stack --------------------------
    int node = kWideEmptyRef;
    if (closest_has_room<E, S>(s)) {
        node = s.cur;
        s.cur = kWideEmptyRef;
        if (node == kWideEmptyRef) {
            while (s.np != base) {
                s.np -= E;
                const uint2 e = lds64(s.np);
                if (__uint_as_float(e.x) < limit) { node = (int)e.y; break; }
            }
            if (node == kWideEmptyRef && s.nspill > 0) {  // deep trees only
                while (s.nspill > 0) {
                    const uint2 e = spill[--s.nspill];
                    if (__uint_as_float(e.x) < limit) { node = (int)e.y; break; }
                }
            }
        }
    } else if (s.tp == ttop) {
        // node stack full and no triangle to drain: move it to the local overflow (rare)
#pragma unroll 1
        for (uint32_t a = base; a != s.np; a += E) spill[s.nspill++] = lds64(a);
        s.np = base;
    }
    if (node == kWideEmptyRef) return;
    if (COUNT) wc->nodes++;
    NodeData n;
    load_node<WIDE>(n, s_nodes, k_smem, sc, node, s.inv.x < 0.f ? 16 : 0, s.inv.y < 0.f ? 48 : 32, s.inv.z < 0.f ? 80 : 64);
    // slab intervals of the four children, packed two per instruction
    const float4 ax = plane_t(n.nx, s.o.x, s.inv.x), bx = plane_t(n.fx, s.o.x, s.inv.x);
    const float4 ay = plane_t(n.ny, s.o.y, s.inv.y), by = plane_t(n.fy, s.o.y, s.inv.y);
    const float4 az = plane_t(n.nz, s.o.z, s.inv.z), bz = plane_t(n.fz, s.o.z, s.inv.z);
    float tn[4], tf[4];
    tn[0] = fmaxf(fmaxf(ax.x, ay.x), fmaxf(az.x, 0.f)); tf[0] = fminf(fminf(bx.x, by.x), fminf(bz.x, limit));
    tn[1] = fmaxf(fmaxf(ax.y, ay.y), fmaxf(az.y, 0.f)); tf[1] = fminf(fminf(bx.y, by.y), fminf(bz.y, limit));
    tn[2] = fmaxf(fmaxf(ax.z, ay.z), fmaxf(az.z, 0.f)); tf[2] = fminf(fminf(bx.z, by.z), fminf(bz.z, limit));
    tn[3] = fmaxf(fmaxf(ax.w, ay.w), fmaxf(az.w, 0.f)); tf[3] = fminf(fminf(bx.w, by.w), fminf(bz.w, limit));
    const int r[4] = {n.ch.x, n.ch.y, n.ch.z, n.ch.w};
    // nearest inner child stays in a register: entry distances are >= 0, so their bit patterns
    // order like the floats; the child slot rides in the two low mantissa bits
    unsigned key[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        key[k] = (tn[k] <= tf[k] && r[k] >= 0) ? ((__float_as_uint(tn[k]) & ~3u) | (unsigned)k) : 0xffffffffu;
    const unsigned best = min(min(key[0], key[1]), min(key[2], key[3]));
    const int r01 = (best & 1u) ? r[1] : r[0], r23 = (best & 1u) ? r[3] : r[2];
    const int rb = (best & 2u) ? r23 : r01;
    s.cur = best != 0xffffffffu ? rb : kWideEmptyRef;
    // big scenes: the node this lane opens next is known now -- start its line on the way to L1 (the step that reads it
    // comes after the other lanes' work of this round; the kernel waits on node fetches more than on anything else there)
    if (WIDE) prefetch_cnode(sc, s.cur, best != 0xffffffffu);
#pragma unroll
    for (int k = 0; k < 4; k++) push_child<E>(s.np, s.tp, tn[k], tf[k], r[k], key[k], best);
}

// Triangle step: the next leaf of the triangle stack that can still matter, TWO of its triangles at once
// (packed FP32, tri_test_pair): most leaves of the SAH tree hold two (C2: 2734 of 3332), so this nearly halves
// the number of triangle steps.  A leaf with one triangle left tests it twice (second result ignored).
template <uint32_t E, int S, bool COUNT>
TRT_DEV void closest_tri_step2(const SceneDev& sc, ClosestRay& s, uint32_t base, WideCounts* wc) {
    const uint32_t ttop = base + (S - 1) * E;
    const float limit = s.d_min * kCullSlack;
    int tri = -1;
    bool two = false;
    while (s.tp != ttop) {
        const uint32_t top = s.tp + E;
        const uint2 e = lds64(top);
        if (!(__uint_as_float(e.x) < limit)) { s.tp = top; continue; }
        const int code = (int)e.y;
        tri = code >> 2;
        const int left = code & 3;  // triangles after this one
        two = left != 0;
        const bool more = left > 1;
        sts64_if(more, top, e.x, (uint32_t)(code + 6));  // first + 2, count - 2
        if (!more) s.tp = top;
        break;
    }
    if (tri < 0) return;
    if (COUNT) wc->tris += two ? 2 : 1;
    const float4* tp = sc.tris + (size_t)tri * 3;
    const float4* tq = tp + (two ? 3 : 0);
    const float4 a0 = __ldg(tp), b0 = __ldg(tp + 1), c0 = __ldg(tp + 2);
    const float4 a1 = __ldg(tq), b1 = __ldg(tq + 1), c1 = __ldg(tq + 2);
    const F3 nd = f3(-s.d.x, -s.d.y, -s.d.z);
    const float2 t = tri_test_pair_records(a0, b0, c0, a1, b1, c1, s.o, s.d, nd);
    const int id0 = f2i(a0.w), id1 = f2i(a1.w);
    if (t.x > 0.f && (t.x < s.d_min || (t.x == s.d_min && (id0 & kTriIdMask) < (s.id & kTriIdMask)))) {
        s.d_min = t.x;  // accepted on the triangle test alone; the winner is verified where it is consumed
        s.id = id0;
    }
    if (two && t.y > 0.f && (t.y < s.d_min || (t.y == s.d_min && (id1 & kTriIdMask) < (s.id & kTriIdMask)))) {
        s.d_min = t.y;
        s.id = id1;
    }
}

template <uint32_t E, int S>
TRT_DEV bool closest_done(const ClosestRay& s, uint32_t base) {
    return s.cur == kWideEmptyRef && s.np == base && s.tp == base + (S - 1) * E && s.nspill == 0;
}

// ---- any hit -----------------------------------------------------------------------------------
struct ShadowRay {
    F3 o, d, inv;
    float max_dist, t_hi;
    uint32_t np, tp;
    int nspill;
    bool occluded;
};

// Stack entries are 32-bit references here (no distances: the interval is fixed); the root is
// pushed at begin.
TRT_DEV void shadow_begin(ShadowRay& s, const float4 o4, const float4 d4, uint32_t base, uint32_t ttop, uint32_t E) {
    s.o = f3(o4.x, o4.y, o4.z);
    s.d = f3(d4.x, d4.y, d4.z);
    s.inv = f3(p_rcp(s.d.x), p_rcp(s.d.y), p_rcp(s.d.z));  // raw reciprocal, reference :276
    s.max_dist = o4.w;
    s.t_hi = p_sub(o4.w, 0.001f);
    s.occluded = false;
    sts32_if(true, base, 0u);  // root
    s.np = base + E;
    s.tp = ttop;
    s.nspill = 0;
}

// TOP PHASE, any hit: returns 1 = occluded by a root-level primitive, 2 = must traverse the tree,
// 0 = unoccluded.  Called by all lanes of a warp.
// A primitive occludes iff its triangle test hits inside the interval AND the reference can reach
// it, i.e. its uploaded leaf box passes the reference's slab test (reference :295-303).  The slab
// test is evaluated FIRST, and only partially: on the box's thinnest axis `a` the products
// t1 = (lo_a - o_a) * inv_a, t2 = (hi_a - o_a) * inv_a are exactly the reference's, and when no
// NaN can occur (all three reciprocals finite: a NaN needs 0 * inf) the reference's
// tmax = min over axes of max(t1, t2) <= max(t1, t2) and tmin >= min(t1, t2), so
// "max(t1,t2) > 0.001 and min(t1,t2) < max_dist" is NECESSARY for the box to pass.  A ray that
// stays inside the room fails it for every wall (the flat wall boxes are 2e-3 thick), at a
// fifth of the cost of the triangle test; the full tests run only for warps where some lane
// passes it.  A warp in which some live lane has a non-finite reciprocal runs the full tests.
// Lanes whose slot holds no shadow ray (`live` false) take no part in the decision.
// one root-level primitive of the shadow top phase; oa / ia are the ray's origin and reciprocal
// direction on the primitive's thin axis
TRT_DEV bool top_shadow_prim(const TopPrims& top, int p, float oa, float ia, const F3 o, const F3 d, const F3 inv,
                             float max_dist, float t_hi, bool take) {
    const float t1 = p_mul(p_sub(top.thin_lo[p], oa), ia), t2 = p_mul(p_sub(top.thin_hi[p], oa), ia);
    const bool maybe = take & (fmaxf(t1, t2) > 0.001f) & (fminf(t1, t2) < max_dist);
    bool occluded = false;
    if (__any_sync(0xffffffffu, maybe)) {
        const float4 a = top.v0[p], b = top.e1[p], c = top.e2[p];
        const float t = tri_test_flat(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), o, d);
        occluded = t > 0.001f && t < t_hi && ref_slab(top.bmin[p], top.bmax[p], o, inv, 0.001f, max_dist);
    }
    return occluded;
}

TRT_DEV int top_shadow(const TopPrims& top, const F3 o, const F3 d, float max_dist, bool live) {
    const F3 inv = f3(p_rcp(d.x), p_rcp(d.y), p_rcp(d.z));  // raw reciprocal, reference :276
    const float t_hi = p_sub(max_dist, 0.001f);
    const float kInf = __int_as_float(0x7f800000);
    const bool finite = fabsf(inv.x) < kInf && fabsf(inv.y) < kInf && fabsf(inv.z) < kInf;
    bool occluded = false;
    if (__all_sync(0xffffffffu, finite | !live)) {
        // the list is sorted by thin axis, so each loop reads the ray's component from a fixed register
        int p = 0;
#pragma unroll 1
        for (const int e = top.n_axis[0]; p < e; p++) occluded |= top_shadow_prim(top, p, o.x, inv.x, o, d, inv, max_dist, t_hi, live);
#pragma unroll 1
        for (const int e = top.n_axis[0] + top.n_axis[1]; p < e; p++) occluded |= top_shadow_prim(top, p, o.y, inv.y, o, d, inv, max_dist, t_hi, live);
#pragma unroll 1
        for (; p < top.n; p++) occluded |= top_shadow_prim(top, p, o.z, inv.z, o, d, inv, max_dist, t_hi, live);
    } else {
        // some lane has an infinite reciprocal (a zero direction component): a NaN could defeat the
        // one-axis argument, so the warp runs the full tests for every primitive
#pragma unroll 1
        for (int p = 0; p < top.n; p++) {
            const float4 a = top.v0[p], b = top.e1[p], c = top.e2[p];
            const float t = tri_test_flat(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), o, d);
            if (t > 0.001f && t < t_hi && ref_slab(top.bmin[p], top.bmax[p], o, inv, 0.001f, max_dist)) occluded = true;
        }
    }
    if (occluded) return 1;
    float tn;
    const bool sx = inv.x < 0.f, sy = inv.y < 0.f, sz = inv.z < 0.f;
    const bool enters = child_interval(sx ? top.root_hi.x : top.root_lo.x, sx ? top.root_lo.x : top.root_hi.x,
                                       sy ? top.root_hi.y : top.root_lo.y, sy ? top.root_lo.y : top.root_hi.y,
                                       sz ? top.root_hi.z : top.root_lo.z, sz ? top.root_lo.z : top.root_hi.z, o, inv,
                                       0.001f, max_dist, &tn);
    return enters ? 2 : 0;
}

template <uint32_t E, int S>
TRT_DEV bool shadow_has_room(const ShadowRay& s) { return (int)(s.tp - s.np) >= (int)(3 * E); }
TRT_DEV bool shadow_node_work(const ShadowRay& s, uint32_t base) {
    return !s.occluded && (s.np != base || s.nspill > 0);
}

template <uint32_t E, int S, bool COUNT, bool WIDE>
TRT_DEV void shadow_node_step(const unsigned char* s_nodes, int k_smem, const SceneDev& sc, ShadowRay& s,
                              uint32_t base, uint32_t* spill, WideCounts* wc) {
    const uint32_t ttop = base + (S - 1) * E;
    int node = kWideEmptyRef;
    if (s.occluded) return;
    if (shadow_has_room<E, S>(s)) {
        if (s.np != base) {
            s.np -= E;
            node = (int)lds32(s.np);
        } else if (s.nspill > 0) {
            node = (int)spill[--s.nspill];
        }
    } else if (s.tp == ttop) {
#pragma unroll 1
        for (uint32_t a = base; a != s.np; a += E) spill[s.nspill++] = lds32(a);
        s.np = base;
    }
    if (node == kWideEmptyRef) return;
    if (COUNT) wc->nodes++;
    NodeData n;
    load_node<WIDE>(n, s_nodes, k_smem, sc, node, s.inv.x < 0.f ? 16 : 0, s.inv.y < 0.f ? 48 : 32, s.inv.z < 0.f ? 80 : 64);
    const float4 ax = plane_t(n.nx, s.o.x, s.inv.x), bx = plane_t(n.fx, s.o.x, s.inv.x);
    const float4 ay = plane_t(n.ny, s.o.y, s.inv.y), by = plane_t(n.fy, s.o.y, s.inv.y);
    const float4 az = plane_t(n.nz, s.o.z, s.inv.z), bz = plane_t(n.fz, s.o.z, s.inv.z);
    // the reference's box interval for shadow rays is (0.001, max_dist)
    const float lo = 0.001f, hi = s.max_dist;
    push_child_any<E>(s.np, s.tp, fmaxf(fmaxf(ax.x, ay.x), fmaxf(az.x, lo)), fminf(fminf(bx.x, by.x), fminf(bz.x, hi)), n.ch.x);
    push_child_any<E>(s.np, s.tp, fmaxf(fmaxf(ax.y, ay.y), fmaxf(az.y, lo)), fminf(fminf(bx.y, by.y), fminf(bz.y, hi)), n.ch.y);
    push_child_any<E>(s.np, s.tp, fmaxf(fmaxf(ax.z, ay.z), fmaxf(az.z, lo)), fminf(fminf(bx.z, by.z), fminf(bz.z, hi)), n.ch.z);
    push_child_any<E>(s.np, s.tp, fmaxf(fmaxf(ax.w, ay.w), fmaxf(az.w, lo)), fminf(fminf(bx.w, by.w), fminf(bz.w, hi)), n.ch.w);
    if (WIDE) {  // the node the next step pops
        const bool any = s.np != base;
        const int nxt = any ? (int)lds32(s.np - E) : 0;
        prefetch_cnode(sc, nxt, any);
    }
}

// one triangle per step (the any-hit work keeps this form: the packed pair test costs registers and a spill in the
// combined kernel and a shadow leaf visit usually ends at its first hit -- measured 98.5 -> 106.7 ms per 64 spp)
template <uint32_t E, int S, bool COUNT>
TRT_DEV void shadow_tri_step(const SceneDev& sc, ShadowRay& s, uint32_t base, WideCounts* wc) {
    const uint32_t ttop = base + (S - 1) * E;
    if (s.occluded || s.tp == ttop) return;
    const uint32_t top = s.tp + E;
    const int scode = (int)lds32(top);
    const int tri = scode >> 2;
    const bool more = (scode & 3) != 0;
    sts32_if(more, top, (uint32_t)(scode + 3));
    if (!more) s.tp = top;
    if (COUNT) wc->tris++;
    const float4* tp = sc.tris + (size_t)tri * 3;
    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
    const F3 v0 = f3(ta.x, ta.y, ta.z), v1 = f3(tb.x, tb.y, tb.z), v2 = f3(tc.x, tc.y, tc.z);
    const float t = tri_test_flat(v0, x_sub(v1, v0), x_sub(v2, v0), s.o, s.d);
    if (t > 0.001f && t < s.t_hi) {
        const int code = f2i(ta.w);
        bool reach;
        if (code & kTriNoDerive) {
            // leaf box not derivable: ask the reference traversal itself (exact, rare)
            Ray r;
            r.o = s.o;
            r.d = s.d;
            VisitCounts vc = {0, 0, 0};
            reach = ref_shadow<false>(sc, r, s.max_dist, &vc);
        } else {
            float4 bmin, bmax;
            derive_leaf_box(v0, v1, v2, &bmin, &bmax);
            reach = ref_slab(bmin, bmax, s.o, s.inv, 0.001f, s.max_dist);
        }
        if (reach) s.occluded = true;
    }
}

// occluder check of one triangle that the test hit inside the interval: the reference must be able to reach it
// (the record is read again here, from L1: keeping six vectors live across the packed test spilled)
TRT_DEV bool shadow_reach(const SceneDev& sc, const ShadowRay& s, const float4* rec) {
    const float4 ta = __ldg(rec), tb = __ldg(rec + 1), tc = __ldg(rec + 2);
    if (f2i(ta.w) & kTriNoDerive) {
        // leaf box not derivable: ask the reference traversal itself (exact, rare)
        Ray r;
        r.o = s.o;
        r.d = s.d;
        VisitCounts vc = {0, 0, 0};
        return ref_shadow<false>(sc, r, s.max_dist, &vc);
    }
    float4 bmin, bmax;
    derive_leaf_box(f3(ta.x, ta.y, ta.z), f3(tb.x, tb.y, tb.z), f3(tc.x, tc.y, tc.z), &bmin, &bmax);
    return ref_slab(bmin, bmax, s.o, s.inv, 0.001f, s.max_dist);
}

// any hit: two triangles of the leaf per step (see closest_tri_step2)
template <uint32_t E, int S, bool COUNT>
TRT_DEV void shadow_tri_step2(const SceneDev& sc, ShadowRay& s, uint32_t base, WideCounts* wc) {
    const uint32_t ttop = base + (S - 1) * E;
    if (s.occluded || s.tp == ttop) return;
    const uint32_t top = s.tp + E;
    const int scode = (int)lds32(top);
    const int tri = scode >> 2;
    const int left = scode & 3;
    const bool two = left != 0, more = left > 1;
    sts32_if(more, top, (uint32_t)(scode + 6));
    if (!more) s.tp = top;
    if (COUNT) wc->tris += two ? 2 : 1;
    const float4* tp = sc.tris + (size_t)tri * 3;
    const float4* tq = tp + (two ? 3 : 0);
    const float4 a0 = __ldg(tp), b0 = __ldg(tp + 1), c0 = __ldg(tp + 2);
    const float4 a1 = __ldg(tq), b1 = __ldg(tq + 1), c1 = __ldg(tq + 2);
    const F3 nd = f3(-s.d.x, -s.d.y, -s.d.z);
    const float2 t = tri_test_pair_records(a0, b0, c0, a1, b1, c1, s.o, s.d, nd);
    const bool h0 = t.x > 0.001f && t.x < s.t_hi, h1 = two && t.y > 0.001f && t.y < s.t_hi;
    if (h0 || h1) {
        bool reach = h0 && shadow_reach(sc, s, tp);
        if (!reach && h1) reach = shadow_reach(sc, s, tq);
        if (reach) s.occluded = true;
    }
}

template <uint32_t E, int S>
TRT_DEV bool shadow_done(const ShadowRay& s, uint32_t base) {
    return s.occluded || (s.np == base && s.tp == base + (S - 1) * E && s.nspill == 0);
}

}  // namespace trt
