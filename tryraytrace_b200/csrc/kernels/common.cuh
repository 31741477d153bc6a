// common.cuh -- device-side types shared by all kernels, and the "pinned" FP32
// operations that parity with the reference depends on.
//
// The reference is compiled with --use_fast_math (reference Makefile:55): FTZ, approximate
// rcp/sqrt/div, and FMA contraction chosen by nvcc/ptxas.  Which multiply-adds were
// contracted decides first-hit ids on silhouette and shared-edge pixels (SURVEY 7.3(1),
// Appendix A.2/A.3; re-derived from the sm_100 SASS of the unmodified kernel).  The
// p_* wrappers below emit exactly one PTX instruction with an explicit rounding mode, so
// ptxas can neither fuse nor split them; the parity-critical code (primary ray, slab test,
// triangle test) is written only in terms of them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trt {

#define TRT_DEV __device__ __forceinline__

TRT_DEV float p_add(float a, float b) { float r; asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
TRT_DEV float p_sub(float a, float b) { float r; asm("sub.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
TRT_DEV float p_mul(float a, float b) { float r; asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
TRT_DEV float p_fma(float a, float b, float c) {
    float r; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
TRT_DEV float p_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
TRT_DEV float p_sqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
TRT_DEV float p_div(float a, float b) { float r; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

struct F3 {
    float x, y, z;
};
TRT_DEV F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }

// Vec operators as the reference kernel's SASS executes them:
//   cross(a,b).x = fma(a.y, b.z, -(a.z*b.y))        first product fused, second rounded
//   dot(a,b)     = fma(a.z,b.z, fma(a.x,b.x, a.y*b.y))
TRT_DEV F3 x_sub(F3 a, F3 b) { return f3(p_sub(a.x, b.x), p_sub(a.y, b.y), p_sub(a.z, b.z)); }
TRT_DEV F3 x_add(F3 a, F3 b) { return f3(p_add(a.x, b.x), p_add(a.y, b.y), p_add(a.z, b.z)); }
TRT_DEV F3 x_scale(F3 a, float s) { return f3(p_mul(a.x, s), p_mul(a.y, s), p_mul(a.z, s)); }
TRT_DEV float x_dot(F3 a, F3 b) { return p_fma(a.z, b.z, p_fma(a.x, b.x, p_mul(a.y, b.y))); }
TRT_DEV F3 x_cross(F3 a, F3 b) {
    return f3(p_fma(a.y, b.z, -p_mul(a.z, b.y)), p_fma(a.z, b.x, -p_mul(a.x, b.z)), p_fma(a.x, b.y, -p_mul(a.y, b.x)));
}
// Vec::norm (reference include/common.h:70-78) as compiled: sqrt.approx of the dot pattern,
// then three multiplies by rcp.approx when the length is positive.
TRT_DEV F3 x_normalize(F3 a) {
    const float len = p_sqrt(p_fma(a.z, a.z, p_fma(a.x, a.x, p_mul(a.y, a.y))));
    if (len > 0.f) {
        const float r = p_rcp(len);
        a = f3(p_mul(a.x, r), p_mul(a.y, r), p_mul(a.z, r));
    }
    return a;
}

struct Ray {
    F3 o, d;
};

// Camera exactly as the 80-byte CameraParams record (reference include/scene.h:64-72).
struct Camera {
    float4 pos, cx, cy, dir;  // .w is padding
    float lens_radius, focus_dist, _p0, _p1;
};
static_assert(sizeof(Camera) == 80, "CameraParams layout");

// Device view of an uploaded scene.
struct SceneDev {
    // reference-layout arrays, byte-identical to what init_scene_data received
    const float4* objects;    // 7 float4 per object (112 B): v0 v1 v2 albedo emission (m,r,ior,tr) (tex_id,pad..)
    const float4* ref_nodes;  // 3 float4 per node (48 B): min, max, (a, b, axis, is_leaf)
    const int* lights;
    int n_objects, n_ref_nodes, n_lights, n_textures;
    cudaTextureObject_t tex[5];
    // re-laid-out arrays for the fast path
    const float4* wide_nodes;  // 8 float4 per 4-wide node (128 B); nullptr when only the compressed form is kept
    const uint4* cnodes;       // compressed form of the same nodes, 4 uint4 each (64 B, traverse_fast.cuh CNode), or nullptr
    const float4* tris;        // 3 float4 per triangle in wide-leaf order: (v0, id | flags) (v1, -) (v2, -)
    int n_wide_nodes, n_tris;
};

// Root-level primitive list (host/wide_bvh.h TopPrim), handed to the traversal kernels BY VALUE:
// it lives in the constant bank, every lane reads the same record, and the brute-force pass over
// it runs fully converged before a ray enters the tree.  root_lo/root_hi bound the tree.
constexpr int kMaxTop = 12;
struct TopPrims {
    float4 v0[kMaxTop];    // v0.xyz, object id | kTriNoDerive (int bits)
    float4 e1[kMaxTop];    // v1 - v0 (single FTZ rounding, as reference :239), unused
    float4 e2[kMaxTop];    // v2 - v0
    float4 bmin[kMaxTop];  // reference leaf box as uploaded
    float4 bmax[kMaxTop];
    float4 root_lo, root_hi;
    // The list is sorted by the axis on which each leaf box is thinnest: primitives [0, n_axis[0]) are
    // thinnest along x, the next n_axis[1] along y, the rest along z; thin_lo / thin_hi are the
    // box planes on that axis (the shadow top phase tests them first, traverse_fast.cuh).
    float thin_lo[kMaxTop], thin_hi[kMaxTop];
    int n_axis[3];
    int n;
    // Closest-hit ranking (traverse_fast.cuh top_rank): inside each axis group the oversized primitives (walls)
    // come first -- n_thin[a] of them, ranked by the thin axis alone; (thin_lo, thin_hi) again as one 8-byte
    // record -- and the small ones (light sources) last; those are listed in full_idx and ranked by the slab
    // interval of their whole leaf box.
    float2 thin2[kMaxTop];
    int n_thin[3];
    int n_full;
    int full_idx[4];
    // Closest-hit brute-force pass, two primitives per instruction (traverse_fast.cuh tri_test_pair): pair j holds
    // the primitives of rank 2j and 2j+1 in ASCENDING OBJECT INDEX (so a tie in t never replaces the earlier one and
    // the update is a plain `<`), component-interleaved: .x = first, .y = second primitive of the pair.  An odd
    // count is padded with a degenerate triangle (zero edges: rejected by the determinant test).
    float2 pair_nv0[kMaxTop / 2][3];  // -v0   (o + (-v0) == o - v0 bit for bit)
    float2 pair_e1[kMaxTop / 2][3];
    float2 pair_ne1[kMaxTop / 2][3];  // -e1   (the reference negates the rounded product; (-a)*b == -(a*b))
    float2 pair_e2[kMaxTop / 2][3];
    int2 pair_id[kMaxTop / 2];        // object id | kTriNoDerive
    int n_pairs;
};

struct RenderConsts {
    int width, height;
    int max_depth, rr_threshold;
};

TRT_DEV int f2i(float f) { return __float_as_int(f); }
TRT_DEV float i2f(int i) { return __int_as_float(i); }

}  // namespace trt
