// bvh.cpp -- CPU BVH builder behind BVH::build (reference src/bvh.cpp:32-113).
//
// Output contract, identical to the reference for the same
// input: median split on the longest axis of the node box, objects ordered by
// centroid with std::sort, one primitive per leaf, 2N-1 nodes in DFS pre-order,
// leaf boxes padded by 1e-3 on axes thinner than 1e-3 (:19-27), inner boxes the
// union of the padded leaf boxes.  The caller's object vector is reordered.
//
// Different from the reference in how it gets there: centroids and padded boxes
// are computed once per object, and each node sorts 8-byte (key, index) proxies
// instead of 112-byte objects.  std::sort's sequence of comparisons and moves
// depends only on comparison outcomes, and the proxy comparator returns exactly
// what the reference comparator returns ((v0+v1+v2)*0.333333f on the split axis,
// :6-9, :85-91), so the resulting permutation -- ties included -- is the same.
#include "bvh.h"
#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {

struct SortKey {
    float key;
    int idx;
};

struct BuildCtx {
    const std::vector<Object>* objs;
    std::vector<AABB> box;       // padded box per original object
    std::vector<Vec> centroid;   // per original object
    std::vector<int> order;      // current permutation
    std::vector<SortKey> scratch;
    std::vector<LinearBVHNode>* nodes;
};

AABB padded_bounds(const Object& o) {
    AABB b = AABB::empty();
    b.grow(o.v0);
    b.grow(o.v1);
    b.grow(o.v2);
    const float pad = 1e-3f;
    const Vec ext = b.max - b.min;
    if (ext.x < pad) { b.min.x -= pad; b.max.x += pad; }
    if (ext.y < pad) { b.min.y -= pad; b.max.y += pad; }
    if (ext.z < pad) { b.min.z -= pad; b.max.z += pad; }
    return b;
}

int build_range(BuildCtx& c, int lo, int hi) {
    const int self = (int)c.nodes->size();
    c.nodes->push_back(LinearBVHNode{});
    std::memset(&(*c.nodes)[self], 0, sizeof(LinearBVHNode));

    AABB bounds = AABB::empty();
    for (int i = lo; i < hi; i++) bounds.grow(c.box[c.order[i]]);
    (*c.nodes)[self].bounds = bounds;

    const int count = hi - lo;
    if (count == 1) {
        LinearBVHNode& n = (*c.nodes)[self];
        n.is_leaf = 1;
        n.primitive_offset = lo;
        n.primitive_count = 1;
        return self;
    }

    const Vec ext = bounds.max - bounds.min;
    int axis = 0;
    if (ext.y > ext.x) axis = 1;
    if (ext.z > ext.y && ext.z > ext.x) axis = 2;
    (*c.nodes)[self].axis = axis;

    SortKey* keys = c.scratch.data() + lo;
    for (int i = lo; i < hi; i++) {
        const int o = c.order[i];
        const Vec& ce = c.centroid[o];
        keys[i - lo] = SortKey{axis == 0 ? ce.x : (axis == 1 ? ce.y : ce.z), o};
    }
    std::sort(keys, keys + count, [](const SortKey& a, const SortKey& b) { return a.key < b.key; });
    for (int i = lo; i < hi; i++) c.order[i] = keys[i - lo].idx;

    const int mid = lo + count / 2;
    const int left = build_range(c, lo, mid);
    const int right = build_range(c, mid, hi);
    LinearBVHNode& n = (*c.nodes)[self];
    n.is_leaf = 0;
    n.left_child_idx = left;
    n.right_child_idx = right;
    return self;
}

}  // namespace

void BVH::build(std::vector<Object>& objects) {
    nodes.clear();
    nodes.reserve(objects.size() * 2);
    if (objects.empty()) return;
    std::printf("[BVH] Building BVH for %lu objects...\n", (unsigned long)objects.size());

    const int n = (int)objects.size();
    BuildCtx c;
    c.objs = &objects;
    c.nodes = &nodes;
    c.box.resize(n);
    c.centroid.resize(n);
    c.order.resize(n);
    c.scratch.resize(n);
    for (int i = 0; i < n; i++) {
        const Object& o = objects[i];
        c.box[i] = padded_bounds(o);
        c.centroid[i] = (o.v0 + o.v1 + o.v2) * 0.333333f;
        c.order[i] = i;
    }
    build_range(c, 0, n);

    std::vector<Object> sorted(n);
    for (int i = 0; i < n; i++) sorted[i] = objects[c.order[i]];
    objects.swap(sorted);
    std::printf("[BVH] Build complete. Total nodes: %lu\n", (unsigned long)nodes.size());
}

// kept for interface parity with the reference class (include/bvh.h:50)
int BVH::build_recursive(std::vector<Object>& objects, int start, int end) {
    (void)objects; (void)start; (void)end;
    return -1;
}
