// wide_bvh.cpp -- see wide_bvh.h.
#include "wide_bvh.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>

namespace trt {
namespace {

struct Box {
    float mn[3], mx[3];
    void reset() {
        for (int k = 0; k < 3; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    }
    void grow(const Box& b) {
        for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], b.mn[k]); mx[k] = std::max(mx[k], b.mx[k]); }
    }
    float area() const {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
    bool valid() const { return mx[0] >= mn[0]; }
};

// binary tree produced by the SAH builder; leaves hold a range of `order`
struct BNode {
    Box box;
    int left = -1, right = -1;  // children (inner) ...
    int first = 0, count = 0;   // ... or primitive range (leaf)
    int item = -1;              // >= 0: a single "top item" leaf (see build_top)
};

constexpr int kBins = 16;
constexpr float kCostBox = 1.0f;  // one child-box test
// leaf policy (tuning knobs TRT_SAH_TRI / TRT_MAX_LEAF; defaults from the B200 sweep)
int g_max_leaf = 4;
float g_cost_tri = 2.5f;  // one triangle test, in child-box tests

struct Builder {
    const std::vector<Box>& pbox;  // per object
    std::vector<int> order;        // objects in the SAH tree, permuted in place
    std::vector<BNode> nodes;

    explicit Builder(const std::vector<Box>& b) : pbox(b) {}

    int build(int first, int count) {
        const int self = (int)nodes.size();
        nodes.emplace_back();
        Box box, cbox;
        box.reset();
        cbox.reset();
        for (int i = first; i < first + count; i++) {
            const Box& b = pbox[order[i]];
            box.grow(b);
            for (int k = 0; k < 3; k++) {
                const float c = 0.5f * (b.mn[k] + b.mx[k]);
                cbox.mn[k] = std::min(cbox.mn[k], c);
                cbox.mx[k] = std::max(cbox.mx[k], c);
            }
        }
        nodes[self].box = box;
        nodes[self].first = first;
        nodes[self].count = count;
        if (count == 1) return self;

        // binned SAH over the three axes
        float best_cost = INFINITY;
        int best_axis = -1, best_split = -1;
        for (int axis = 0; axis < 3; axis++) {
            const float ext = cbox.mx[axis] - cbox.mn[axis];
            if (!(ext > 0.f)) continue;
            Box bb[kBins];
            int bc[kBins];
            for (int b = 0; b < kBins; b++) { bb[b].reset(); bc[b] = 0; }
            const float scale = kBins / ext;
            for (int i = first; i < first + count; i++) {
                const Box& pb = pbox[order[i]];
                int b = (int)((0.5f * (pb.mn[axis] + pb.mx[axis]) - cbox.mn[axis]) * scale);
                b = std::min(std::max(b, 0), kBins - 1);
                bb[b].grow(pb);
                bc[b]++;
            }
            float right_area[kBins];
            int right_cnt[kBins];
            Box acc;
            acc.reset();
            int cnt = 0;
            for (int b = kBins - 1; b > 0; b--) {
                acc.grow(bb[b]);
                cnt += bc[b];
                right_area[b] = acc.area();
                right_cnt[b] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int b = 0; b < kBins - 1; b++) {
                acc.grow(bb[b]);
                cnt += bc[b];
                if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                const float cost = acc.area() * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
            }
        }
        const float leaf_cost = g_cost_tri * count;
        const float split_cost =
            best_axis < 0 ? INFINITY : 2.f * kCostBox + g_cost_tri * best_cost / std::max(box.area(), 1e-30f);
        if (count <= g_max_leaf && leaf_cost <= split_cost) return self;

        int mid;
        if (best_axis < 0) {
            mid = first + count / 2;  // all centroids coincide: split the range in half
        } else {
            const float ext = cbox.mx[best_axis] - cbox.mn[best_axis];
            const float scale = kBins / ext;
            const float cmin = cbox.mn[best_axis];
            const int axis = best_axis, split = best_split;
            int* b0 = order.data() + first;
            int* m = std::partition(b0, b0 + count, [&](int o) {
                const Box& pb = pbox[o];
                int b = (int)((0.5f * (pb.mn[axis] + pb.mx[axis]) - cmin) * scale);
                b = std::min(std::max(b, 0), kBins - 1);
                return b <= split;
            });
            mid = (int)(m - order.data());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        const int l = build(first, mid - first);
        const int r = build(mid, first + count - mid);
        nodes[self].left = l;
        nodes[self].right = r;
        nodes[self].count = 0;
        return self;
    }
};

// An unused child slot gets an inverted infinite box (lo = +inf, hi = -inf): whichever plane the
// ray's direction sign selects as "near", the slab interval comes out empty, so the traversal
// kernels need no special case for it.
void clear_node(WideNode& w) {
    std::memset(&w, 0, sizeof(w));
    for (int k = 0; k < 4; k++) {
        w.lo_x[k] = w.lo_y[k] = w.lo_z[k] = INFINITY;
        w.hi_x[k] = w.hi_y[k] = w.hi_z[k] = -INFINITY;
        w.child[k] = kWideEmpty;
    }
}

// single rounding, denormal inputs and results flushed like the device's FADD.FTZ
float ftz(float r) { return std::fabs(r) < 1.17549435e-38f ? std::copysign(0.f, r) : r; }
float add_ftz(float a, float b) { return ftz(ftz(a) + ftz(b)); }

// The leaf box the traversal kernels derive from a triangle's vertices (kernels/traverse_fast.cuh
// derive_leaf_box; the rule of reference src/bvh.cpp:12-30), evaluated the way the device does.
void derived_leaf_box(const Object& o, float mn[3], float mx[3]) {
    const float v[3][3] = {{o.v0.x, o.v0.y, o.v0.z}, {o.v1.x, o.v1.y, o.v1.z}, {o.v2.x, o.v2.y, o.v2.z}};
    for (int k = 0; k < 3; k++) {
        float lo = ftz(std::fmin(std::fmin(v[0][k], v[1][k]), v[2][k]));  // FMNMX.FTZ on the device
        float hi = ftz(std::fmax(std::fmax(v[0][k], v[1][k]), v[2][k]));
        if (add_ftz(hi, -lo) < 1e-3f) { lo = add_ftz(lo, -1e-3f); hi = add_ftz(hi, 1e-3f); }
        mn[k] = lo;
        mx[k] = hi;
    }
}

}  // namespace

void build_wide_bvh(const Object* objects, int n_objects, const LinearBVHNode* ref_nodes, int n_ref_nodes,
                    WideBvh& out) {
    if (const char* e = getenv("TRT_SAH_TRI")) g_cost_tri = std::max(0.1f, (float)atof(e));
    if (const char* e = getenv("TRT_MAX_LEAF")) g_max_leaf = std::max(1, std::min(4, atoi(e)));
    out.nodes.clear();
    out.tris.clear();
    out.leaf_boxes.assign(n_objects, LeafBox{{INFINITY, INFINITY, INFINITY, 0}, {-INFINITY, -INFINITY, -INFINITY, 0}});
    out.n_top_prims = 0;
    out.n_underivable = 0;
    out.depth = 0;

    // 1. reference leaf box of every object (objects no leaf refers to can never be hit)
    std::vector<Box> pbox(n_objects);
    for (auto& b : pbox) b.reset();
    for (int i = 0; i < n_ref_nodes; i++) {
        const LinearBVHNode& n = ref_nodes[i];
        if (!n.is_leaf) continue;
        for (int k = 0; k < n.primitive_count; k++) {
            const int o = n.primitive_offset + k;
            Box b;
            b.mn[0] = n.bounds.min.x; b.mn[1] = n.bounds.min.y; b.mn[2] = n.bounds.min.z;
            b.mx[0] = n.bounds.max.x; b.mx[1] = n.bounds.max.y; b.mx[2] = n.bounds.max.z;
            if (pbox[o].valid()) pbox[o].grow(b); else pbox[o] = b;
        }
    }
    std::vector<int> live;
    Box scene;
    scene.reset();
    for (int o = 0; o < n_objects; o++) {
        if (!pbox[o].valid()) continue;
        live.push_back(o);
        scene.grow(pbox[o]);
        LeafBox& lb = out.leaf_boxes[o];
        for (int k = 0; k < 3; k++) { lb.mn[k] = pbox[o].mn[k]; lb.mx[k] = pbox[o].mx[k]; }
    }
    out.top.clear();
    for (int k = 0; k < 3; k++) { out.root_mn[k] = INFINITY; out.root_mx[k] = -INFINITY; }
    if (live.empty()) {
        WideNode n;
        clear_node(n);
        out.nodes.push_back(n);
        return;
    }

    // 2. lift oversized primitives out of the SAH tree
    const float scene_area = std::max(scene.area(), 1e-30f);
    std::vector<int> top, rest;
    {
        std::vector<int> by_area = live;
        std::sort(by_area.begin(), by_area.end(), [&](int a, int b) {
            const float aa = pbox[a].area(), ab = pbox[b].area();
            return aa != ab ? aa > ab : a < b;
        });
        const size_t max_top = kMaxTopPrims;
        std::vector<char> is_top(n_objects, 0);
        if (live.size() > 16)
            for (size_t i = 0; i < by_area.size() && top.size() < max_top; i++) {
                if (pbox[by_area[i]].area() < 0.02f * scene_area) break;
                top.push_back(by_area[i]);
                is_top[by_area[i]] = 1;
            }
        // Light sources (emission above the reference's listing threshold, src/main.cpp:93) join the
        // root-level list when there are only a few of them: every shadow ray ends just short of a
        // light, so a light inside the tree stretches the tree's bounding box over all of them and
        // sends every shadow ray (and most closest-hit rays) into the tree for nothing.
        if (live.size() > 16) {
            std::vector<int> emitters;
            for (int o : live) {
                const Object& ob = objects[o];
                if (!is_top[o] && (ob.emission.x > 0.1f || ob.emission.y > 0.1f || ob.emission.z > 0.1f)) emitters.push_back(o);
            }
            if (emitters.size() <= (size_t)kMaxTopLights && top.size() + emitters.size() <= max_top)
                for (int o : emitters) { top.push_back(o); is_top[o] = 1; }
        }
        for (int o : live)
            if (!is_top[o]) rest.push_back(o);
        if (rest.empty()) { rest = live; top.clear(); }
    }
    out.n_top_prims = (int)top.size();

    // 3. binned-SAH binary tree over the rest
    Builder bld(pbox);
    bld.order = rest;
    const int mesh_root = bld.build(0, (int)rest.size());

    // 4. the root-level list
    const int root = mesh_root;
    const std::vector<int>& order = bld.order;
    out.top.clear();
    for (int o : top) {
        const Object& ob = objects[o];
        TopPrim tp;
        tp.v0[0] = ob.v0.x; tp.v0[1] = ob.v0.y; tp.v0[2] = ob.v0.z;
        {   // like the tree's records: flag a primitive whose uploaded leaf box the vertex rule does not reproduce
            float dmn[3], dmx[3];
            derived_leaf_box(ob, dmn, dmx);
            bool same = true;
            for (int k = 0; k < 3; k++) same = same && dmn[k] == pbox[o].mn[k] && dmx[k] == pbox[o].mx[k];
            tp.id = o | (same ? 0 : kTriNoDeriveBit);
            if (!same) out.n_underivable++;
        }
        tp.e1[0] = add_ftz(ob.v1.x, -ob.v0.x); tp.e1[1] = add_ftz(ob.v1.y, -ob.v0.y); tp.e1[2] = add_ftz(ob.v1.z, -ob.v0.z);
        tp.e2[0] = add_ftz(ob.v2.x, -ob.v0.x); tp.e2[1] = add_ftz(ob.v2.y, -ob.v0.y); tp.e2[2] = add_ftz(ob.v2.z, -ob.v0.z);
        for (int k = 0; k < 3; k++) { tp.mn[k] = pbox[o].mn[k]; tp.mx[k] = pbox[o].mx[k]; }
        out.top.push_back(tp);
    }
    for (int k = 0; k < 3; k++) { out.root_mn[k] = bld.nodes[root].box.mn[k]; out.root_mx[k] = bld.nodes[root].box.mx[k]; }

    // 5. triangle records in leaf order
    out.tris.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) {
        const Object& o = objects[order[i]];
        TriRecord& t = out.tris[i];
        t.v0[0] = o.v0.x; t.v0[1] = o.v0.y; t.v0[2] = o.v0.z;
        t.v1[0] = o.v1.x; t.v1[1] = o.v1.y; t.v1[2] = o.v1.z;
        t.v2[0] = o.v2.x; t.v2[1] = o.v2.y; t.v2[2] = o.v2.z;
        t.pad1 = t.pad2 = 0.f;
        // the kernels re-derive the reference leaf box from the vertices; where that does not
        // reproduce the uploaded leaf node exactly (a foreign builder, NaNs), flag the triangle
        float mn[3], mx[3];
        derived_leaf_box(o, mn, mx);
        const Box& lb = pbox[order[i]];
        bool same = true;
        for (int k = 0; k < 3; k++) same = same && mn[k] == lb.mn[k] && mx[k] == lb.mx[k];
        t.id = order[i] | (same ? 0 : kTriNoDeriveBit);
        if (!same) out.n_underivable++;
    }

    // 6. collapse the binary tree into 4-wide nodes
    const std::vector<BNode>& bn = bld.nodes;
    auto is_leaf = [&](int n) { return bn[n].left < 0; };
    auto leaf_ref = [&](int n) { return ~((bn[n].first << 2) | (bn[n].count - 1)); };
    if (is_leaf(root)) {  // a single leaf: wrap it in one node
        WideNode w;
        clear_node(w);
        w.lo_x[0] = bn[root].box.mn[0]; w.hi_x[0] = bn[root].box.mx[0];
        w.lo_y[0] = bn[root].box.mn[1]; w.hi_y[0] = bn[root].box.mx[1];
        w.lo_z[0] = bn[root].box.mn[2]; w.hi_z[0] = bn[root].box.mx[2];
        w.child[0] = leaf_ref(root);
        out.nodes.push_back(w);
        out.depth = 1;
        return;
    }
    // Nodes are emitted in descending box area (a proxy for how often rays visit them), so the
    // first k nodes are the ones worth staging into shared memory.
    struct Work { int bnode; int wide; int depth; float area; };
    auto colder = [](const Work& a, const Work& b) { return a.area < b.area; };
    std::vector<Work> work;  // max-heap on area
    out.nodes.emplace_back();
    work.push_back(Work{root, 0, 1, bn[root].box.area()});
    while (!work.empty()) {
        std::pop_heap(work.begin(), work.end(), colder);
        const Work wk = work.back();
        work.pop_back();
        out.depth = std::max(out.depth, wk.depth);
        int kids[4];
        int nk = 0;
        kids[nk++] = bn[wk.bnode].left;
        kids[nk++] = bn[wk.bnode].right;
        while (nk < 4) {  // open the inner child with the largest box
            int pick = -1;
            float pa = -1.f;
            for (int k = 0; k < nk; k++)
                if (!is_leaf(kids[k]) && bn[kids[k]].box.area() > pa) { pa = bn[kids[k]].box.area(); pick = k; }
            if (pick < 0) break;
            const int n = kids[pick];
            kids[pick] = bn[n].left;
            kids[nk++] = bn[n].right;
        }
        WideNode w;
        clear_node(w);
        for (int k = 0; k < 4; k++) {
            if (k >= nk) continue;
            const Box& b = bn[kids[k]].box;
            w.lo_x[k] = b.mn[0]; w.hi_x[k] = b.mx[0];
            w.lo_y[k] = b.mn[1]; w.hi_y[k] = b.mx[1];
            w.lo_z[k] = b.mn[2]; w.hi_z[k] = b.mx[2];
            if (is_leaf(kids[k])) {
                w.child[k] = leaf_ref(kids[k]);
            } else {
                const int idx = (int)out.nodes.size();
                out.nodes.emplace_back();
                w.child[k] = idx;
                work.push_back(Work{kids[k], idx, wk.depth + 1, b.area()});
                std::push_heap(work.begin(), work.end(), colder);
            }
        }
        out.nodes[wk.wide] = w;
    }
}

}  // namespace trt
