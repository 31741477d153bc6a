// bvh_build.cuh -- device-side construction of the 4-wide BVH (LBVH, Karras 2012) from the
// uploaded object array.  Replaces, for large scenes, the host re-layout of csrc/host/wide_bvh.cpp
// (which itself replaces what the traversal reads of reference src/bvh.cpp:32-113): the
// reference builder is a single-threaded recursive median split (34 s for 10 M triangles,
// SURVEY section 6); this one is a handful of kernels over the objects already in HBM.
//
// The output has exactly the layout and invariants wide_bvh.h documents (WideNode, TriRecord,
// TopPrim): every wide box contains the reference leaf boxes below it, oversized primitives are
// lifted into the root-level list, leaves hold 1..4 triangles that are contiguous in the triangle
// record array, records carry kTriNoDeriveBit when the vertex rule does not reproduce the uploaded
// leaf box.  So the exactness argument of traverse_fast.cuh holds unchanged for either builder.
#pragma once
#include "common.cuh"
#include <cuda_runtime.h>
#include <string>

namespace trt {

struct DeviceWideBvh {
    float4* d_nodes = nullptr;  // n_nodes x 128 B (cudaMallocAsync on `s`: free with cudaFreeAsync; owned by the caller)
    float4* d_tris = nullptr;   // n_tris x 48 B
    int n_nodes = 0, n_tris = 0;
    TopPrims top{};             // root-level list + tree bounds, ready to hand to the kernels
    int n_top = 0;
    int n_underivable = 0;
    int depth = 0;
    float build_ms = 0.f;       // device time, CUDA events on `s`
};

// d_objects: n_objects x 112 B (7 float4).  d_ref_nodes: the reference node array (3 float4 per
// node) whose leaves define each object's reference leaf box, or nullptr: then the box is derived
// from the vertices with the reference builder's rule (reference src/bvh.cpp:12-30).
// max_leaf in [1, 4].  top_sah: rebuild the upper levels with a binned-SAH pass over LBVH clusters
// (HLBVH-style hybrid; a few thousand boxes, done on the host).  Returns 0, or -1 with *err set.
int build_wide_bvh_device(const float4* d_objects, int n_objects, const float4* d_ref_nodes, int n_ref_nodes,
                          int max_leaf, bool top_sah, DeviceWideBvh* out, cudaStream_t s, std::string* err);

}  // namespace trt
