// wide_bvh.h -- re-layout of the reference BVH output for the fast traversal path.
//
// Input: the object array (already in BVH::build order) and the reference's 48-byte node
// array (reference include/bvh.h:12-28, built by src/bvh.cpp:32-113).  Output:
//   * a 4-wide BVH (128-byte nodes, child boxes stored SoA so every plane set is one
//     128-bit load) built by binned SAH over the REFERENCE LEAF BOXES -- so every wide box
//     contains the reference leaf boxes below it, which is what makes the fast traversal a
//     strict superset of the reference traversal (kernels/traverse_wide.cuh);
//   * oversized primitives (the room's one-triangle walls) lifted out of
//     the SAH tree into a short root-level list (`top`, at most kMaxTopPrims): the traversal
//     kernels test that list brute force, fully converged, before a ray ever enters the tree,
//     so the giant boxes neither inflate the upper levels nor cost stack traffic;
//   * 48-byte triangle records (v0, v1, v2 as three float4) in wide-leaf order; the kernels
//     subtract the edges per test exactly as the reference does, and re-derive the reference
//     leaf box (bounds of the three vertices + the builder's padding rule) to decide exactly
//     whether the reference traversal would have reached a candidate triangle.  A triangle
//     whose uploaded leaf box is NOT reproduced by that rule carries kTriNoDeriveBit in its id
//     word and is always resolved by the reference-order replay;
//   * the reference leaf box of every object (host side only: builder input, tests).
#pragma once
#include "bvh.h"
#include "scene.h"
#include <cstdint>
#include <vector>

namespace trt {

constexpr int kWideEmpty = 0x7fffffff;  // child slot unused

struct WideNode {  // 128 bytes
    float lo_x[4], hi_x[4], lo_y[4], hi_y[4], lo_z[4], hi_z[4];
    int child[4];  // >= 0 inner node index; < 0 leaf: ~((first_tri << 2) | (count - 1)); kWideEmpty unused
    int pad[4];
};
static_assert(sizeof(WideNode) == 128, "WideNode layout");

constexpr int kTriNoDeriveBit = 0x40000000;

struct TriRecord {  // 48 bytes
    float v0[3];
    int id;  // object index (position in the reference-sorted array) | kTriNoDeriveBit
    float v1[3];
    float pad1;
    float v2[3];
    float pad2;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord layout");

struct LeafBox {  // 32 bytes
    float mn[4], mx[4];
};

constexpr int kMaxTopPrims = 12;
constexpr int kMaxTopLights = 4;  // light sources lifted into the root-level list, when there are at most this many

// A root-level primitive: vertices as (v0, e1 = v1 - v0, e2 = v2 - v0) with the single FTZ
// rounding the reference applies per test (reference src/renderer.cu:239-240), and its
// reference leaf box as uploaded.
struct TopPrim {
    float v0[3];
    int id;
    float e1[3], e2[3];
    float mn[3], mx[3];
};

struct WideBvh {
    std::vector<WideNode> nodes;  // root = 0; the tree covers every object that is not in `top`
    std::vector<TopPrim> top;     // root-level list (oversized primitives)
    float root_mn[3] = {0, 0, 0}, root_mx[3] = {0, 0, 0};  // bounds of the tree (inverted when empty)
    std::vector<TriRecord> tris;
    std::vector<LeafBox> leaf_boxes;  // per object id
    int n_top_prims = 0;
    int n_underivable = 0;  // triangles flagged kTriNoDeriveBit
    int depth = 0;
};

void build_wide_bvh(const Object* objects, int n_objects, const LinearBVHNode* ref_nodes, int n_ref_nodes,
                    WideBvh& out);

}  // namespace trt
