// bvh_build.cu -- see bvh_build.cuh.  LBVH over 63-bit Morton codes of the reference leaf-box
// centroids, sorted by a hand-written stable LSD radix sort (Karras 2012: one thread per internal node finds its key range and split with
// count-leading-zeros searches; boxes are fitted bottom-up with one atomic counter per node),
// then a breadth-first collapse into 4-wide nodes (open the inner child with the largest box,
// like the host builder) with device-side allocation of the output nodes.
#include "bvh_build.cuh"
#include "traverse_fast.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace trt {
namespace {

constexpr int kB = 256;
constexpr int kFlagValid = 1, kFlagTop = 2;  // kTriNoDerive (bit 30) rides in the same word
constexpr int kCandCap = 1024;
constexpr int kLightCap = 4;  // host/wide_bvh.h kMaxTopLights

struct Bounds {
    float mn[3], mx[3];
    int count, aux;
};

struct Task {
    int bnode, wide, depth;
};

__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}

__global__ void k_init_bounds(Bounds* b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inf = __int_as_float(0x7f800000);
    for (int k = 0; k < 3; k++) { b[i].mn[k] = inf; b[i].mx[k] = -inf; }
    b[i].count = b[i].aux = 0;
}

__global__ void k_init_boxes(float4* lo, float4* hi, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inf = __int_as_float(0x7f800000);
    lo[i] = make_float4(inf, inf, inf, 0.f);
    hi[i] = make_float4(-inf, -inf, -inf, 0.f);
}

// reference leaf box of every object = union of the leaf nodes that refer to it
__global__ void k_boxes_from_ref(const float4* __restrict__ ref_nodes, int n_nodes, float4* lo, float4* hi,
                                 int n_objects) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int4 link = __ldg(reinterpret_cast<const int4*>(ref_nodes + (size_t)i * 3 + 2));
    if (!link.w) return;
    const float4 bmin = __ldg(ref_nodes + (size_t)i * 3), bmax = __ldg(ref_nodes + (size_t)i * 3 + 1);
    for (int k = 0; k < link.y; k++) {
        const int o = link.x + k;
        if (o < 0 || o >= n_objects) continue;
        atomic_min_f(&lo[o].x, bmin.x); atomic_min_f(&lo[o].y, bmin.y); atomic_min_f(&lo[o].z, bmin.z);
        atomic_max_f(&hi[o].x, bmax.x); atomic_max_f(&hi[o].y, bmax.y); atomic_max_f(&hi[o].z, bmax.z);
    }
}

__global__ void k_boxes_derived(const float4* __restrict__ objects, int n, float4* lo, float4* hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* op = objects + (size_t)i * 7;
    const float4 a = __ldg(op), b = __ldg(op + 1), c = __ldg(op + 2);
    float4 bmin, bmax;
    derive_leaf_box(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), &bmin, &bmax);
    lo[i] = bmin;
    hi[i] = bmax;
}

// block-wide min/max of a box + count, one set of atomics per block
__device__ void reduce_bounds(bool take, float3 mn, float3 mx, Bounds* out) {
    __shared__ float s_mn[3][kB / 32], s_mx[3][kB / 32];
    __shared__ int s_cnt[kB / 32];
    const float inf = __int_as_float(0x7f800000);
    float v[6] = {take ? mn.x : inf, take ? mn.y : inf, take ? mn.z : inf, take ? mx.x : -inf, take ? mx.y : -inf,
                  take ? mx.z : -inf};
    int cnt = take ? 1 : 0;
    for (int off = 16; off; off >>= 1) {
        for (int k = 0; k < 3; k++) {
            v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], off));
            v[3 + k] = fmaxf(v[3 + k], __shfl_xor_sync(0xffffffffu, v[3 + k], off));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        for (int k = 0; k < 3; k++) { s_mn[k][warp] = v[k]; s_mx[k][warp] = v[3 + k]; }
        s_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        float m[6] = {inf, inf, inf, -inf, -inf, -inf};
        for (int w = 0; w < kB / 32; w++) {
            for (int k = 0; k < 3; k++) { m[k] = fminf(m[k], s_mn[k][w]); m[3 + k] = fmaxf(m[3 + k], s_mx[k][w]); }
            total += s_cnt[w];
        }
        if (total) {
            for (int k = 0; k < 3; k++) { atomic_min_f(&out->mn[k], m[k]); atomic_max_f(&out->mx[k], m[3 + k]); }
            atomicAdd(&out->count, total);
        }
    }
}

// validity + "the vertex rule reproduces the uploaded leaf box" flag per object, scene bounds
__global__ void __launch_bounds__(kB) k_flags(const float4* __restrict__ objects, int n, const float4* __restrict__ lo,
                                              const float4* __restrict__ hi, int* flags, Bounds* scene) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    float4 l = make_float4(0, 0, 0, 0), h = l;
    if (i < n) {
        l = lo[i];
        h = hi[i];
        valid = h.x >= l.x && h.y >= l.y && h.z >= l.z;
        int f = 0;
        if (valid) {
            const float4* op = objects + (size_t)i * 7;
            const float4 a = __ldg(op), b = __ldg(op + 1), c = __ldg(op + 2);
            float4 bmin, bmax;
            derive_leaf_box(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), &bmin, &bmax);
            const bool same = bmin.x == l.x && bmin.y == l.y && bmin.z == l.z && bmax.x == h.x && bmax.y == h.y && bmax.z == h.z;
            f = kFlagValid | (same ? 0 : kTriNoDerive);
            if (!same) atomicAdd(&scene->aux, 1);
        }
        flags[i] = f;
    }
    reduce_bounds(valid, make_float3(l.x, l.y, l.z), make_float3(h.x, h.y, h.z), scene);
}

__device__ __forceinline__ float box_area(float4 l, float4 h) {
    const float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    if (dx < 0.f || dy < 0.f || dz < 0.f) return 0.f;
    return 2.f * (dx * dy + dy * dz + dz * dx);
}

// candidates for the root-level list: oversized boxes, and light sources (emission above the
// reference's listing threshold, src/main.cpp:93; see host/wide_bvh.cpp for why)
__global__ void k_select_top(const float4* __restrict__ objects, const float4* __restrict__ lo,
                             const float4* __restrict__ hi, const int* __restrict__ flags, int n, float thr_area,
                             int2* cand, int* n_cand, int* lights, int* n_lights) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !(flags[i] & kFlagValid)) return;
    const float a = box_area(lo[i], hi[i]);
    if (a >= thr_area) {
        const int k = atomicAdd(n_cand, 1);
        if (k < kCandCap) cand[k] = make_int2(i, __float_as_int(a));
        return;
    }
    const float4 e = __ldg(objects + (size_t)i * 7 + 4);
    if (e.x > 0.1f || e.y > 0.1f || e.z > 0.1f) {
        const int k = atomicAdd(n_lights, 1);
        if (k < kLightCap) lights[k] = i;
    }
}

__global__ void k_mark_top(int* flags, const int* ids, int n_top) {
    const int i = threadIdx.x;
    if (i < n_top) flags[ids[i]] |= kFlagTop;
}

__global__ void __launch_bounds__(kB) k_centroid_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                        const int* __restrict__ flags, int n, Bounds* cb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    float3 c = make_float3(0, 0, 0);
    if (i < n) {
        const int f = flags[i];
        take = (f & kFlagValid) && !(f & kFlagTop);
        const float4 l = lo[i], h = hi[i];
        c = make_float3(0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z));
    }
    reduce_bounds(take, c, c, cb);
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ flags,
                         int n, float3 cmin, float3 scale, unsigned long long* keys, int* idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = i;
    const int f = flags[i];
    if (!(f & kFlagValid) || (f & kFlagTop)) {
        keys[i] = ~0ull;  // sorts behind every real key
        return;
    }
    const float4 l = lo[i], h = hi[i];
    const float q[3] = {(0.5f * (l.x + h.x) - cmin.x) * scale.x, (0.5f * (l.y + h.y) - cmin.y) * scale.y,
                        (0.5f * (l.z + h.z) - cmin.z) * scale.z};
    unsigned long long k = 0;
    for (int a = 0; a < 3; a++) {
        const float v = fminf(fmaxf(q[a], 0.f), 2097151.f);
        k |= spread21((unsigned long long)v) << (2 - a);
    }
    keys[i] = k;  // 63 bits: strictly below ~0ull
}

// ---- stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass ---------------------
// Three kernels per pass.  k_radix_hist: every block counts the digits of its contiguous tile.
// k_radix_scan: one block turns the (digit-major, block-minor) count table into global offsets.
// k_radix_scatter: every block walks its tile again in order, 256 keys at a time; a key's rank
// among the equal digits of its chunk comes from __match_any_sync inside the warp plus per-warp
// digit counts combined across the eight warps in warp order, which keeps the pass stable.
constexpr int kRadixBits = 8, kRadixBins = 1 << kRadixBits, kRadixBlock = 256;

__global__ void __launch_bounds__(kRadixBlock) k_radix_hist(const unsigned long long* __restrict__ keys, int n, int tile,
                                                            int shift, int* __restrict__ table, int n_blocks) {
    __shared__ int hist[kRadixBins];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int begin = blockIdx.x * tile, end = min(n, begin + tile);
    for (int i = begin + threadIdx.x; i < end; i += kRadixBlock) atomicAdd(&hist[(int)((keys[i] >> shift) & (kRadixBins - 1))], 1);
    __syncthreads();
    table[threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of `count` ints in place, one block of 1024 threads
__global__ void __launch_bounds__(1024) k_radix_scan(int* table, int count) {
    __shared__ int part[1024];
    const int per = (count + 1023) / 1024;
    const int begin = threadIdx.x * per, end = min(count, begin + per);
    int sum = 0;
    for (int i = begin; i < end; i++) sum += table[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the partial sums
        const int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (int i = begin; i < end; i++) {
        const int c = table[i];
        table[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(kRadixBlock) k_radix_scatter(const unsigned long long* __restrict__ keys_in,
                                                               const int* __restrict__ vals_in, unsigned long long* keys_out,
                                                               int* vals_out, int n, int tile, int shift,
                                                               const int* __restrict__ table, int n_blocks) {
    __shared__ int base[kRadixBins];                    // next output slot of each digit for this block
    __shared__ int wcount[kRadixBlock / 32][kRadixBins];  // per-warp digit counts of the current chunk
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    base[threadIdx.x] = table[threadIdx.x * n_blocks + blockIdx.x];
    const int begin = blockIdx.x * tile, end = min(n, begin + tile);
    for (int chunk = begin; chunk < end; chunk += kRadixBlock) {
        for (int w = 0; w < kRadixBlock / 32; w++) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        const int i = chunk + threadIdx.x;
        const bool on = i < end;
        unsigned long long key = 0;
        int val = 0, digit = 0, rank = 0;
        if (on) {
            key = keys_in[i];
            val = vals_in[i];
            digit = (int)((key >> shift) & (kRadixBins - 1));
        }
        // lanes past the end get a digit of their own class so that they never join a real group
        const unsigned peers = __match_any_sync(0xffffffffu, on ? digit : kRadixBins + (int)lane);
        if (on) {
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) wcount[warp][digit] = __popc(peers);
        }
        __syncthreads();
        {   // thread d: offsets of digit d for the eight warps, in warp order
            int run = base[threadIdx.x];
            for (int w = 0; w < kRadixBlock / 32; w++) {
                const int c = wcount[w][threadIdx.x];
                wcount[w][threadIdx.x] = run;
                run += c;
            }
            base[threadIdx.x] = run;
        }
        __syncthreads();
        if (on) {
            const int dst = wcount[warp][digit] + rank;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, one thread per internal node.  Children: >= 0 internal node, < 0 leaf ~index
// (index into the sorted order).
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int* left, int* right, int* parent_int,
                            int* parent_leaf, int* first, int* last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int a = min(i, j), b = max(i, j);
    const int lc = a == gamma ? ~gamma : gamma;
    const int rc = b == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    first[i] = a;
    last[i] = b;
    if (lc < 0) parent_leaf[~lc] = i; else parent_int[lc] = i;
    if (rc < 0) parent_leaf[~rc] = i; else parent_int[rc] = i;
    if (i == 0) parent_int[0] = -1;
}

__global__ void k_fit(int n_leaves, const int* __restrict__ sorted, const float4* __restrict__ olo,
                      const float4* __restrict__ ohi, const int* __restrict__ parent_leaf,
                      const int* __restrict__ parent_int, const int* __restrict__ left, const int* __restrict__ right,
                      float4* ilo, float4* ihi, int* visits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    int cur = parent_leaf[i];
    while (cur >= 0) {
        if (atomicAdd(&visits[cur], 1) == 0) return;  // the sibling subtree is not finished yet
        __threadfence();
        float4 l[2], h[2];
        const int c[2] = {left[cur], right[cur]};
        for (int k = 0; k < 2; k++) {
            if (c[k] < 0) {
                const int o = sorted[~c[k]];
                l[k] = olo[o];
                h[k] = ohi[o];
            } else {
                l[k] = __ldcg(&ilo[c[k]]);
                h[k] = __ldcg(&ihi[c[k]]);
            }
        }
        ilo[cur] = make_float4(fminf(l[0].x, l[1].x), fminf(l[0].y, l[1].y), fminf(l[0].z, l[1].z), 0.f);
        ihi[cur] = make_float4(fmaxf(h[0].x, h[1].x), fmaxf(h[0].y, h[1].y), fmaxf(h[0].z, h[1].z), 0.f);
        __threadfence();
        cur = parent_int[cur];
    }
}

// Cluster roots of the LBVH: the subtrees of at most `cmax` primitives whose parent is larger.
// The host builds a SAH tree over their boxes (the upper levels, where Morton splits are poorest);
// below a cluster root the LBVH is kept.  ref >= 0 internal node, < 0 leaf ~index.
struct Cluster {
    float4 lo, hi;  // lo.w carries the node reference (int bits)
};
__global__ void k_clusters(int n_leaves, int cmax, const int* __restrict__ first, const int* __restrict__ last,
                           const int* __restrict__ parent_int, const int* __restrict__ parent_leaf,
                           const int* __restrict__ sorted, const float4* __restrict__ olo, const float4* __restrict__ ohi,
                           const float4* __restrict__ ilo, const float4* __restrict__ ihi, Cluster* out, int cap, int* n_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_int = n_leaves - 1;
    if (i >= n_int + n_leaves) return;
    int ref, parent, count;
    if (i < n_int) {
        ref = i;
        parent = parent_int[i];
        count = last[i] - first[i] + 1;
    } else {
        ref = ~(i - n_int);
        parent = parent_leaf[i - n_int];
        count = 1;
    }
    if (count > cmax || parent < 0) return;
    if (last[parent] - first[parent] + 1 <= cmax) return;  // the parent is inside a cluster already
    const int k = atomicAdd(n_out, 1);
    if (k >= cap) return;
    float4 l, h;
    if (ref >= 0) { l = ilo[ref]; h = ihi[ref]; } else { const int o = sorted[~ref]; l = olo[o]; h = ohi[o]; }
    out[k].lo = make_float4(l.x, l.y, l.z, __int_as_float(ref));
    out[k].hi = h;
}

struct TreeView {
    const int *left, *right, *first, *last, *sorted;
    const float4 *ilo, *ihi, *olo, *ohi;
    int max_leaf;
};

__device__ __forceinline__ void child_box(const TreeView& t, int c, float4* l, float4* h);

// A subtree of at most max_leaf primitives becomes ONE leaf when that is cheaper than opening it
// by the host builder's SAH rule (host/wide_bvh.cpp: a triangle test costs 2.5 child-box tests):
//     2.5 * n  <=  2 + 2.5 * (A_l * n_l + A_r * n_r) / A
__device__ __forceinline__ bool leaf_like(const TreeView& t, int c) {
    if (c < 0) return true;
    const int n = t.last[c] - t.first[c] + 1;
    if (n > t.max_leaf) return false;
    const float a = box_area(t.ilo[c], t.ihi[c]);
    float split = 0.f;
    const int kid[2] = {t.left[c], t.right[c]};
    for (int k = 0; k < 2; k++) {
        float4 l, h;
        child_box(t, kid[k], &l, &h);
        const int nk = kid[k] < 0 ? 1 : t.last[kid[k]] - t.first[kid[k]] + 1;
        split += box_area(l, h) * (float)nk;
    }
    return 2.5f * (float)n <= 2.f + 2.5f * split / fmaxf(a, 1e-30f);
}
__device__ __forceinline__ void child_box(const TreeView& t, int c, float4* l, float4* h) {
    if (c < 0) {
        const int o = t.sorted[~c];
        *l = t.olo[o];
        *h = t.ohi[o];
    } else {
        *l = t.ilo[c];
        *h = t.ihi[c];
    }
}

__device__ void write_wide(float4* nodes, int w, int nk, const float4* l, const float4* h, const int* ref) {
    const float inf = __int_as_float(0x7f800000);
    float v[6][4];
    int ch[4];
    for (int k = 0; k < 4; k++) {
        const bool on = k < nk;
        v[0][k] = on ? l[k].x : inf; v[1][k] = on ? h[k].x : -inf;
        v[2][k] = on ? l[k].y : inf; v[3][k] = on ? h[k].y : -inf;
        v[4][k] = on ? l[k].z : inf; v[5][k] = on ? h[k].z : -inf;
        ch[k] = on ? ref[k] : kWideEmptyRef;
    }
    float4* p = nodes + (size_t)w * 8;
    for (int r = 0; r < 6; r++) p[r] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    p[6] = make_float4(__int_as_float(ch[0]), __int_as_float(ch[1]), __int_as_float(ch[2]), __int_as_float(ch[3]));
    p[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// one thread per wide node of the current level
__global__ void k_collapse(const Task* __restrict__ in, int n_in, Task* out, int* n_out, int* n_wide, int* max_depth,
                           TreeView t, float4* nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const Task task = in[i];
    int kids[4];
    int nk = 2;
    kids[0] = t.left[task.bnode];
    kids[1] = t.right[task.bnode];
    while (nk < 4) {  // open the inner child with the largest box
        int pick = -1;
        float pa = -1.f;
        for (int k = 0; k < nk; k++) {
            if (leaf_like(t, kids[k])) continue;
            const float a = box_area(t.ilo[kids[k]], t.ihi[kids[k]]);
            if (a > pa) { pa = a; pick = k; }
        }
        if (pick < 0) break;
        const int c = kids[pick];
        kids[pick] = t.left[c];
        kids[nk++] = t.right[c];
    }
    float4 l[4], h[4];
    int ref[4];
    for (int k = 0; k < nk; k++) {
        const int c = kids[k];
        child_box(t, c, &l[k], &h[k]);
        if (leaf_like(t, c)) {
            const int f = c < 0 ? ~c : t.first[c];
            const int cnt = c < 0 ? 1 : t.last[c] - t.first[c] + 1;
            ref[k] = ~((f << 2) | (cnt - 1));
        } else {
            const int w = atomicAdd(n_wide, 1);
            ref[k] = w;
            const int o = atomicAdd(n_out, 1);
            out[o] = Task{c, w, task.depth + 1};
        }
    }
    write_wide(nodes, task.wide, nk, l, h, ref);
    atomicMax(max_depth, task.depth);
}

// the whole tree is one leaf: a single node with one child
__global__ void k_single_leaf(TreeView t, int n_rest, float4* nodes, float4* root_lo, float4* root_hi) {
    if (threadIdx.x || blockIdx.x) return;
    const float inf = __int_as_float(0x7f800000);
    float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
    for (int i = 0; i < n_rest; i++) {
        const int o = t.sorted[i];
        const float4 a = t.olo[o], b = t.ohi[o];
        l = make_float4(fminf(l.x, a.x), fminf(l.y, a.y), fminf(l.z, a.z), 0.f);
        h = make_float4(fmaxf(h.x, b.x), fmaxf(h.y, b.y), fmaxf(h.z, b.z), 0.f);
    }
    const int ref = ~((0 << 2) | (n_rest - 1));
    write_wide(nodes, 0, 1, &l, &h, &ref);
    *root_lo = l;
    *root_hi = h;
}

__global__ void k_tri_records(const float4* __restrict__ objects, const int* __restrict__ sorted,
                              const int* __restrict__ flags, int n_rest, float4* tris) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rest) return;
    const int o = sorted[i];
    const float4* op = objects + (size_t)o * 7;
    const float4 a = __ldg(op), b = __ldg(op + 1), c = __ldg(op + 2);
    float4* t = tris + (size_t)i * 3;
    t[0] = make_float4(a.x, a.y, a.z, __int_as_float(o | (flags[o] & kTriNoDerive)));
    t[1] = make_float4(b.x, b.y, b.z, 0.f);
    t[2] = make_float4(c.x, c.y, c.z, 0.f);
}

int grid(long long n) { return (int)((n + kB - 1) / kB); }
float __int_as_float_host(int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

// host mirror of the device's FADD.FTZ (single rounding, denormals flushed)
float ftz(float r) { return std::fabs(r) < 1.17549435e-38f ? std::copysign(0.f, r) : r; }
float sub_ftz(float a, float b) { return ftz(ftz(a) - ftz(b)); }

// Binned SAH (16 bins) over the cluster boxes, one cluster per leaf.  Output: binary nodes with
// children >= 0 = top node index, < 0 = ~cluster index; node 0 is the root.
struct TopNode {
    float lo[3], hi[3];
    int left, right;
};
struct TopSah {
    const std::vector<Cluster>& c;
    std::vector<int> order;
    std::vector<TopNode> nodes;
    explicit TopSah(const std::vector<Cluster>& cl) : c(cl), order(cl.size()) {
        for (size_t i = 0; i < cl.size(); i++) order[i] = (int)i;
        nodes.reserve(cl.size());
    }
    static float area(const float* lo, const float* hi) {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx < 0 || dy < 0 || dz < 0) ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
    }
    static void grow(float* lo, float* hi, const Cluster& k) {
        lo[0] = std::min(lo[0], k.lo.x); lo[1] = std::min(lo[1], k.lo.y); lo[2] = std::min(lo[2], k.lo.z);
        hi[0] = std::max(hi[0], k.hi.x); hi[1] = std::max(hi[1], k.hi.y); hi[2] = std::max(hi[2], k.hi.z);
    }
    static float centroid(const Cluster& k, int a) {
        return a == 0 ? 0.5f * (k.lo.x + k.hi.x) : (a == 1 ? 0.5f * (k.lo.y + k.hi.y) : 0.5f * (k.lo.z + k.hi.z));
    }
    // returns the child reference of the subtree over order[first, first + count)
    int build(int first, int count) {
        if (count == 1) return ~order[first];
        const int self = (int)nodes.size();
        nodes.emplace_back();
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = first; i < first + count; i++) {
            const Cluster& k = c[order[i]];
            grow(lo, hi, k);
            for (int a = 0; a < 3; a++) { const float x = centroid(k, a); clo[a] = std::min(clo[a], x); chi[a] = std::max(chi[a], x); }
        }
        constexpr int B = 16;
        float best = INFINITY;
        int best_axis = -1, best_split = -1;
        for (int a = 0; a < 3; a++) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.f)) continue;
            float blo[B][3], bhi[B][3];
            int bc[B];
            for (int b = 0; b < B; b++) { for (int k = 0; k < 3; k++) { blo[b][k] = INFINITY; bhi[b][k] = -INFINITY; } bc[b] = 0; }
            const float scale = B / ext;
            for (int i = first; i < first + count; i++) {
                const Cluster& k = c[order[i]];
                const int b = std::min(std::max((int)((centroid(k, a) - clo[a]) * scale), 0), B - 1);
                grow(blo[b], bhi[b], k);
                bc[b]++;
            }
            float ra[B];
            int rc[B];
            float alo[3] = {INFINITY, INFINITY, INFINITY}, ahi[3] = {-INFINITY, -INFINITY, -INFINITY};
            int cnt = 0;
            for (int b = B - 1; b > 0; b--) {
                for (int k = 0; k < 3; k++) { alo[k] = std::min(alo[k], blo[b][k]); ahi[k] = std::max(ahi[k], bhi[b][k]); }
                cnt += bc[b];
                ra[b] = area(alo, ahi);
                rc[b] = cnt;
            }
            for (int k = 0; k < 3; k++) { alo[k] = INFINITY; ahi[k] = -INFINITY; }
            cnt = 0;
            for (int b = 0; b < B - 1; b++) {
                for (int k = 0; k < 3; k++) { alo[k] = std::min(alo[k], blo[b][k]); ahi[k] = std::max(ahi[k], bhi[b][k]); }
                cnt += bc[b];
                if (cnt == 0 || rc[b + 1] == 0) continue;
                const float cost = area(alo, ahi) * cnt + ra[b + 1] * rc[b + 1];
                if (cost < best) { best = cost; best_axis = a; best_split = b; }
            }
        }
        int mid = first + count / 2;
        if (best_axis >= 0) {
            const float ext = chi[best_axis] - clo[best_axis], scale = B / ext, cmin = clo[best_axis];
            const int a = best_axis, split = best_split;
            int* b0 = order.data() + first;
            int* m = std::partition(b0, b0 + count, [&](int o) {
                return std::min(std::max((int)((centroid(c[o], a) - cmin) * scale), 0), B - 1) <= split;
            });
            const int mm = (int)(m - order.data());
            if (mm != first && mm != first + count) mid = mm;
        }
        const int l = build(first, mid - first);
        const int r = build(mid, first + count - mid);
        TopNode& n = nodes[self];
        for (int k = 0; k < 3; k++) { n.lo[k] = lo[k]; n.hi[k] = hi[k]; }
        n.left = l;
        n.right = r;
        return self;
    }
};

// Scratch buffers come from the device's stream-ordered memory pool: the build needs ~300 bytes per
// primitive for a few milliseconds, and cudaMalloc / cudaFree of gigabytes (with their device-wide
// synchronisation and page mapping) used to cost more than the kernels; the pool keeps the memory
// for the next build of the process.  Everything is released when the object goes out of scope.
struct Scratch {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : stream(s) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;  // do not hand the memory back to the driver between builds
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, stream);
    }
    template <class T>
    cudaError_t get(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(count, 1) * sizeof(T), stream);
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = (T*)q;
        return e;
    }
};

}  // namespace

#define BCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            if (err) *err = std::string(#call) + " failed: " + cudaGetErrorString(e_);                  \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

int build_wide_bvh_device(const float4* d_objects, int n, const float4* d_ref_nodes, int n_ref_nodes, int max_leaf,
                          bool top_sah, DeviceWideBvh* out, cudaStream_t s, std::string* err) {
    max_leaf = std::max(1, std::min(4, max_leaf));
    *out = DeviceWideBvh();
    Scratch tmp(s);
    struct Events {  // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } evs;
    BCU(cudaEventCreate(&evs.a));
    BCU(cudaEventCreate(&evs.b));
    const cudaEvent_t ev0 = evs.a, ev1 = evs.b;
    BCU(cudaEventRecord(ev0, s));

    float4 *olo, *ohi;
    int* flags;
    Bounds* d_bounds;  // [0] scene, [1] centroids of the tree's objects
    BCU(tmp.get(&olo, n));
    BCU(tmp.get(&ohi, n));
    BCU(tmp.get(&flags, n));
    BCU(tmp.get(&d_bounds, 2));
    k_init_bounds<<<1, 32, 0, s>>>(d_bounds, 2);

    // 1. reference leaf box per object
    if (d_ref_nodes && n_ref_nodes > 0) {
        k_init_boxes<<<grid(n), kB, 0, s>>>(olo, ohi, n);
        k_boxes_from_ref<<<grid(n_ref_nodes), kB, 0, s>>>(d_ref_nodes, n_ref_nodes, olo, ohi, n);
    } else {
        k_boxes_derived<<<grid(n), kB, 0, s>>>(d_objects, n, olo, ohi);
    }
    k_flags<<<grid(n), kB, 0, s>>>(d_objects, n, olo, ohi, flags, d_bounds);
    Bounds hb[2];
    BCU(cudaMemcpyAsync(hb, d_bounds, sizeof(Bounds), cudaMemcpyDeviceToHost, s));
    BCU(cudaStreamSynchronize(s));
    const int n_live = hb[0].count;
    out->n_underivable = hb[0].aux;
    const float inf = INFINITY;
    out->top.root_lo = make_float4(inf, inf, inf, 0.f);
    out->top.root_hi = make_float4(-inf, -inf, -inf, 0.f);
    if (n_live == 0) {  // nothing can be hit: one empty node
        BCU(cudaMallocAsync(&out->d_nodes, 128, s));
        BCU(cudaMallocAsync(&out->d_tris, 48, s));
        const float4 empty[8] = {make_float4(inf, inf, inf, inf), make_float4(-inf, -inf, -inf, -inf),
                                 make_float4(inf, inf, inf, inf), make_float4(-inf, -inf, -inf, -inf),
                                 make_float4(inf, inf, inf, inf), make_float4(-inf, -inf, -inf, -inf),
                                 make_float4(__int_as_float_host(kWideEmptyRef), __int_as_float_host(kWideEmptyRef),
                                             __int_as_float_host(kWideEmptyRef), __int_as_float_host(kWideEmptyRef)),
                                 make_float4(0, 0, 0, 0)};
        BCU(cudaMemcpyAsync(out->d_nodes, empty, 128, cudaMemcpyHostToDevice, s));
        BCU(cudaStreamSynchronize(s));
        out->n_nodes = 1;
        return 0;
    }

    // 2. oversized primitives -> root-level list (same rule as the host builder: box area above
    //    2 % of the scene box, the 12 largest, only for scenes of more than 16 primitives)
    std::vector<int> top_ids;
    {
        const float dx = hb[0].mx[0] - hb[0].mn[0], dy = hb[0].mx[1] - hb[0].mn[1], dz = hb[0].mx[2] - hb[0].mn[2];
        const float scene_area = std::max(2.f * (dx * dy + dy * dz + dz * dx), 1e-30f);
        if (n_live > 16) {
            int2* d_cand;
            int *d_ncand, *d_lights;  // d_ncand[0] oversized candidates, d_ncand[1] light sources
            BCU(tmp.get(&d_cand, kCandCap));
            BCU(tmp.get(&d_ncand, 2));
            BCU(tmp.get(&d_lights, kLightCap));
            BCU(cudaMemsetAsync(d_ncand, 0, 8, s));
            k_select_top<<<grid(n), kB, 0, s>>>(d_objects, olo, ohi, flags, n, 0.02f * scene_area, d_cand, d_ncand,
                                                d_lights, d_ncand + 1);
            int h_n[2] = {0, 0};
            BCU(cudaMemcpyAsync(h_n, d_ncand, 8, cudaMemcpyDeviceToHost, s));
            BCU(cudaStreamSynchronize(s));
            const int n_cand = h_n[0], n_emit = h_n[1];
            if (n_cand > 0 && n_cand <= kCandCap) {
                std::vector<int2> cand(n_cand);
                BCU(cudaMemcpy(cand.data(), d_cand, sizeof(int2) * n_cand, cudaMemcpyDeviceToHost));
                std::sort(cand.begin(), cand.end(), [](const int2& a, const int2& b) {
                    float fa, fb;
                    memcpy(&fa, &a.y, 4);
                    memcpy(&fb, &b.y, 4);
                    return fa != fb ? fa > fb : a.x < b.x;
                });
                for (int i = 0; i < n_cand && (int)top_ids.size() < kMaxTop; i++) top_ids.push_back(cand[i].x);
            }
            if (n_emit > 0 && n_emit <= kLightCap && (int)top_ids.size() + n_emit <= kMaxTop) {
                int ids[kLightCap];
                BCU(cudaMemcpy(ids, d_lights, 4 * n_emit, cudaMemcpyDeviceToHost));
                std::sort(ids, ids + n_emit);
                for (int i = 0; i < n_emit; i++) top_ids.push_back(ids[i]);
            }
            if ((int)top_ids.size() == n_live) top_ids.clear();
        }
    }
    const int n_top = (int)top_ids.size();
    if (n_top) {
        int* d_ids;
        BCU(tmp.get(&d_ids, kMaxTop));
        BCU(cudaMemcpyAsync(d_ids, top_ids.data(), 4 * n_top, cudaMemcpyHostToDevice, s));
        k_mark_top<<<1, 32, 0, s>>>(flags, d_ids, n_top);
    }
    out->n_top = n_top;

    // 3. Morton order of the tree's objects
    k_centroid_bounds<<<grid(n), kB, 0, s>>>(olo, ohi, flags, n, d_bounds + 1);
    BCU(cudaMemcpyAsync(hb + 1, d_bounds + 1, sizeof(Bounds), cudaMemcpyDeviceToHost, s));
    BCU(cudaStreamSynchronize(s));
    const int n_rest = hb[1].count;
    float3 cmin = make_float3(hb[1].mn[0], hb[1].mn[1], hb[1].mn[2]);
    float3 scale;
    {
        const float e[3] = {hb[1].mx[0] - hb[1].mn[0], hb[1].mx[1] - hb[1].mn[1], hb[1].mx[2] - hb[1].mn[2]};
        scale = make_float3(e[0] > 0.f ? 2097152.f / e[0] : 0.f, e[1] > 0.f ? 2097152.f / e[1] : 0.f,
                            e[2] > 0.f ? 2097152.f / e[2] : 0.f);
    }
    unsigned long long *keys_a, *keys_b;
    int *idx_a, *idx_b;
    BCU(tmp.get(&keys_a, n));
    BCU(tmp.get(&keys_b, n));
    BCU(tmp.get(&idx_a, n));
    BCU(tmp.get(&idx_b, n));
    k_morton<<<grid(n), kB, 0, s>>>(olo, ohi, flags, n, cmin, scale, keys_a, idx_a);
    {   // 63-bit keys (and the all-ones key of excluded objects): eight 8-bit passes, ping-pong a <-> b
        const int n_blocks = std::max(1, std::min(1024, (n + 4095) / 4096));
        const int tile = ((n + n_blocks - 1) / n_blocks + kRadixBlock - 1) / kRadixBlock * kRadixBlock;
        int* table;
        BCU(tmp.get(&table, (size_t)kRadixBins * n_blocks));
        unsigned long long *kin = keys_a, *kout = keys_b;
        int *vin = idx_a, *vout = idx_b;
        for (int pass = 0; pass < 8; pass++) {
            const int shift = pass * kRadixBits;
            k_radix_hist<<<n_blocks, kRadixBlock, 0, s>>>(kin, n, tile, shift, table, n_blocks);
            k_radix_scan<<<1, 1024, 0, s>>>(table, kRadixBins * n_blocks);
            k_radix_scatter<<<n_blocks, kRadixBlock, 0, s>>>(kin, vin, kout, vout, n, tile, shift, table, n_blocks);
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        // an even number of passes: the sorted data is back in keys_a / idx_a; keep the names used below
        std::swap(keys_a, keys_b);
        std::swap(idx_a, idx_b);
    }
    const unsigned long long* keys = keys_b;
    const int* sorted = idx_b;  // the first n_rest entries are the tree's objects in Morton order

    // 4. triangle records in leaf order
    BCU(cudaMallocAsync(&out->d_tris, (size_t)std::max(n_rest, 1) * 48, s));
    out->n_tris = n_rest;
    k_tri_records<<<grid(n_rest), kB, 0, s>>>(d_objects, sorted, flags, n_rest, out->d_tris);

    // 5. hierarchy, boxes, collapse
    TreeView tv;
    memset(&tv, 0, sizeof(tv));
    tv.sorted = sorted;
    tv.olo = olo;
    tv.ohi = ohi;
    tv.max_leaf = max_leaf;
    float4* d_root;  // root_lo, root_hi
    BCU(tmp.get(&d_root, 2));
    if (n_rest <= max_leaf) {
        BCU(cudaMallocAsync(&out->d_nodes, 128, s));
        k_single_leaf<<<1, 32, 0, s>>>(tv, n_rest, out->d_nodes, d_root, d_root + 1);
        out->n_nodes = 1;
        out->depth = 1;
    } else {
        const int n_int = n_rest - 1;
        const int cluster_cap = 65536;  // > 2 * 8192 cluster roots plus stragglers
        const int top_cap = cluster_cap;
        int *left, *right, *parent_int, *parent_leaf, *first, *last, *visits;
        float4 *ilo, *ihi;
        BCU(tmp.get(&left, n_int + top_cap));
        BCU(tmp.get(&right, n_int + top_cap));
        BCU(tmp.get(&parent_int, n_int));
        BCU(tmp.get(&parent_leaf, n_rest));
        BCU(tmp.get(&first, n_int + top_cap));
        BCU(tmp.get(&last, n_int + top_cap));
        BCU(tmp.get(&visits, n_int));
        BCU(tmp.get(&ilo, n_int + top_cap));
        BCU(tmp.get(&ihi, n_int + top_cap));
        BCU(cudaMemsetAsync(visits, 0, (size_t)n_int * 4, s));
        k_hierarchy<<<grid(n_int), kB, 0, s>>>(keys, n_rest, left, right, parent_int, parent_leaf, first, last);
        k_fit<<<grid(n_rest), kB, 0, s>>>(n_rest, sorted, olo, ohi, parent_leaf, parent_int, left, right, ilo, ihi, visits);
        tv.left = left; tv.right = right; tv.first = first; tv.last = last; tv.ilo = ilo; tv.ihi = ihi;

        // SAH over clusters for the upper levels (HLBVH-style hybrid): cut the LBVH into subtrees of
        // at most cmax primitives, build a binned-SAH tree over their boxes on the host (a few
        // thousand boxes), and splice it in as extra internal nodes [n_int, n_int + n_top) above them
        int root_ref = 0;
        const int cmax = std::max(2 * max_leaf, n_rest / 8192);
        if (top_sah && n_rest > 4 * cmax) {
            BCU(cudaStreamSynchronize(s));
            std::vector<Cluster> cl(cluster_cap);
            Cluster* d_cl;
            int* d_ncl;
            BCU(tmp.get(&d_cl, cluster_cap));
            BCU(tmp.get(&d_ncl, 1));
            BCU(cudaMemsetAsync(d_ncl, 0, 4, s));
            k_clusters<<<grid(n_int + n_rest), kB, 0, s>>>(n_rest, cmax, first, last, parent_int, parent_leaf, sorted, olo,
                                                           ohi, ilo, ihi, d_cl, cluster_cap, d_ncl);
            int n_cl = 0;
            BCU(cudaMemcpyAsync(&n_cl, d_ncl, 4, cudaMemcpyDeviceToHost, s));
            BCU(cudaStreamSynchronize(s));
            if (n_cl > 1 && n_cl <= cluster_cap) {
                cl.resize(n_cl);
                BCU(cudaMemcpy(cl.data(), d_cl, sizeof(Cluster) * n_cl, cudaMemcpyDeviceToHost));
                // device-side append order is arbitrary: sort by node reference so the tree is reproducible
                std::sort(cl.begin(), cl.end(), [](const Cluster& a, const Cluster& b) {
                    int ra, rb;
                    memcpy(&ra, &a.lo.w, 4);
                    memcpy(&rb, &b.lo.w, 4);
                    return ra < rb;
                });
                TopSah sah(cl);
                sah.build(0, n_cl);
                const int n_top = (int)sah.nodes.size();  // = n_cl - 1 <= top_cap
                std::vector<int> hl(n_top), hr(n_top), hf(n_top, 0), hla(n_top, 0x3fffffff);
                std::vector<float4> hlo(n_top), hhi(n_top);
                auto conv = [&](int c) {
                    if (c >= 0) return n_int + c;
                    int ref;
                    memcpy(&ref, &cl[~c].lo.w, 4);
                    return ref;
                };
                for (int i = 0; i < n_top; i++) {
                    const TopNode& t = sah.nodes[i];
                    hl[i] = conv(t.left);
                    hr[i] = conv(t.right);
                    hlo[i] = make_float4(t.lo[0], t.lo[1], t.lo[2], 0.f);
                    hhi[i] = make_float4(t.hi[0], t.hi[1], t.hi[2], 0.f);
                }
                BCU(cudaMemcpyAsync(left + n_int, hl.data(), 4 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaMemcpyAsync(right + n_int, hr.data(), 4 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaMemcpyAsync(first + n_int, hf.data(), 4 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaMemcpyAsync(last + n_int, hla.data(), 4 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaMemcpyAsync(ilo + n_int, hlo.data(), 16 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaMemcpyAsync(ihi + n_int, hhi.data(), 16 * (size_t)n_top, cudaMemcpyHostToDevice, s));
                BCU(cudaStreamSynchronize(s));  // the host vectors go out of scope
                root_ref = n_int;
            }
        }

        float4* wide_tmp;  // a wide node consumes at least one internal node
        Task *fr_a, *fr_b;
        int* d_cnt;  // [0] next-level tasks, [1] wide nodes allocated, [2] depth
        BCU(tmp.get(&wide_tmp, (size_t)(n_int + top_cap) * 8));
        BCU(tmp.get(&fr_a, n_int + top_cap));
        BCU(tmp.get(&fr_b, n_int + top_cap));
        BCU(tmp.get(&d_cnt, 4));
        const Task root_task = {root_ref, 0, 1};
        const int init_cnt[4] = {0, 1, 0, 0};
        BCU(cudaMemcpyAsync(fr_a, &root_task, sizeof(Task), cudaMemcpyHostToDevice, s));
        BCU(cudaMemcpyAsync(d_cnt, init_cnt, sizeof(init_cnt), cudaMemcpyHostToDevice, s));
        int n_in = 1;
        int h_cnt[4];
        for (int level = 0; n_in > 0; level++) {
            if (level > 512) {
                if (err) *err = "device BVH collapse did not terminate";
                return -1;
            }
            k_collapse<<<grid(n_in), kB, 0, s>>>(fr_a, n_in, fr_b, d_cnt, d_cnt + 1, d_cnt + 2, tv, wide_tmp);
            BCU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, s));
            BCU(cudaStreamSynchronize(s));
            n_in = h_cnt[0];
            BCU(cudaMemsetAsync(d_cnt, 0, 4, s));
            std::swap(fr_a, fr_b);
        }
        out->n_nodes = h_cnt[1];
        out->depth = h_cnt[2];
        BCU(cudaMallocAsync(&out->d_nodes, (size_t)out->n_nodes * 128, s));
        BCU(cudaMemcpyAsync(out->d_nodes, wide_tmp, (size_t)out->n_nodes * 128, cudaMemcpyDeviceToDevice, s));
        BCU(cudaMemcpyAsync(d_root, ilo + root_ref, 16, cudaMemcpyDeviceToDevice, s));
        BCU(cudaMemcpyAsync(d_root + 1, ihi + root_ref, 16, cudaMemcpyDeviceToDevice, s));
    }
    float4 h_root[2];
    BCU(cudaMemcpyAsync(h_root, d_root, 32, cudaMemcpyDeviceToHost, s));

    // 6. the root-level list, on the host (at most 12 primitives)
    TopPrims& tp = out->top;
    memset(&tp, 0, sizeof(tp));
    tp.n = n_top;
    for (int i = 0; i < n_top; i++) {
        float4 ob[3], bl, bh;
        int f;
        BCU(cudaMemcpyAsync(ob, d_objects + (size_t)top_ids[i] * 7, 48, cudaMemcpyDeviceToHost, s));
        BCU(cudaMemcpyAsync(&bl, olo + top_ids[i], 16, cudaMemcpyDeviceToHost, s));
        BCU(cudaMemcpyAsync(&bh, ohi + top_ids[i], 16, cudaMemcpyDeviceToHost, s));
        BCU(cudaMemcpyAsync(&f, flags + top_ids[i], 4, cudaMemcpyDeviceToHost, s));
        BCU(cudaStreamSynchronize(s));
        tp.v0[i] = make_float4(ob[0].x, ob[0].y, ob[0].z, __int_as_float_host(top_ids[i] | (f & kTriNoDerive)));
        tp.e1[i] = make_float4(sub_ftz(ob[1].x, ob[0].x), sub_ftz(ob[1].y, ob[0].y), sub_ftz(ob[1].z, ob[0].z), 0.f);
        tp.e2[i] = make_float4(sub_ftz(ob[2].x, ob[0].x), sub_ftz(ob[2].y, ob[0].y), sub_ftz(ob[2].z, ob[0].z), 0.f);
        tp.bmin[i] = make_float4(bl.x, bl.y, bl.z, 0.f);
        tp.bmax[i] = make_float4(bh.x, bh.y, bh.z, 0.f);
    }
    BCU(cudaEventRecord(ev1, s));
    BCU(cudaStreamSynchronize(s));
    tp.root_lo = make_float4(h_root[0].x, h_root[0].y, h_root[0].z, 0.f);
    tp.root_hi = make_float4(h_root[1].x, h_root[1].y, h_root[1].z, 0.f);
    BCU(cudaEventElapsedTime(&out->build_ms, ev0, ev1));
    BCU(cudaGetLastError());
    return 0;
}

}  // namespace trt
