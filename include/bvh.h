// bvh.h -- linear BVH node + CPU builder.
// Drop-in for the reference's include/bvh.h: LinearBVHNode (:12-28, 48 bytes)
// and class BVH (:33-51).  The builder reproduces src/bvh.cpp:32-113 (median
// split on the longest axis, one primitive per leaf, DFS pre-order array).
#pragma once
#include "common.h"
#include "scene.h"
#include "aabb.h"
#include <vector>

struct __align__(16) LinearBVHNode {
    AABB bounds;
    union {
        int left_child_idx;    // inner: left child (always self + 1)
        int primitive_offset;  // leaf: first object
    };
    union {
        int right_child_idx;   // inner
        int primitive_count;   // leaf
    };
    int axis;     // split axis of an inner node
    int is_leaf;  // 1 = leaf
};

static_assert(sizeof(LinearBVHNode) == 48, "LinearBVHNode layout");

class BVH {
public:
    // NOTE: reorders `objects` -- every object index handed to the renderer
    // (hit ids, light indices) refers to the sorted array.
    void build(std::vector<Object>& objects);
    const std::vector<LinearBVHNode>& get_nodes() const { return nodes; }

private:
    std::vector<LinearBVHNode> nodes;
    int build_recursive(std::vector<Object>& objects, int start, int end);
};
