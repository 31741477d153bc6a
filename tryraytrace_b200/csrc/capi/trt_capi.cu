// trt_capi.cu -- the C ABI (include/trt_capi.h): context, scene upload and re-layout,
// the wavefront render loop, parity entry points, and C-callable forms of the host surface.
#include "trt_capi.h"

#include "bvh.h"
#include "camera.h"
#include "image_io.h"
#include "loader.h"
#include "scene.h"

#include "../host/wide_bvh.h"
#include "../host/xorwow_tables.h"
#include "../kernels/bvh_build.cuh"
#include "../kernels/wavefront.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace trt;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(TRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kFrameChunkDefault = 256;  // frames per wavefront job (bounds the column-vector table); TRT_FRAME_CHUNK overrides
constexpr int kBatchIterations = 16;  // iterations issued between completion polls
constexpr int kTailBatchIterations = 4;  // ... once the job is within reach of its drain phase: the poll drives the grid sizes there
constexpr int kWideStackEntries = 128;  // kernels/traverse_fast.cuh kSpillEntries
constexpr int kAutoDeviceBuildAbove = 1 << 18;  // TRT_BUILD_AUTO: objects above which the device builder is used

}  // namespace

struct trt_ctx {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // created by trt_create
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_poll[2] = {nullptr, nullptr};
    float last_ms = 0.f;
    unsigned long long launches = 0;

    // scene
    bool have_scene = false;
    float4* d_objects = nullptr;
    float4* d_ref_nodes = nullptr;
    int* d_lights = nullptr;
    float4* d_wide_nodes = nullptr;
    uint4* d_cnodes = nullptr;  // compressed 64-byte form of the wide nodes (big scenes; then d_wide_nodes is released)
    float4* d_tris = nullptr;
    bool wide_from_pool = false;  // d_wide_nodes / d_tris came from the stream-ordered pool (device builder)
    std::vector<cudaArray_t> tex_arrays;
    std::vector<cudaTextureObject_t> tex_objs;
    SceneDev sc{};
    TopPrims top{};
    trt_scene_info info{};

    // XORWOW tables
    int rng_w = 0, rng_h = 0;
    uint32_t* d_row_a = nullptr;  // window tables of the row matrices (words 0-3 / word 4)
    uint32_t* d_row_b = nullptr;
    uint32_t* d_col_a = nullptr;  // two-level column window tables (host/xorwow_tables.h xorwow_build_col_levels): words 0-3 ...
    uint32_t* d_col_b = nullptr;  // ... and word 4 of the entries
    XwColVec* d_col_vecs = nullptr;
    size_t col_vecs_cap = 0;

    // wavefront pool
    int pool_cap = 0;
    void* pool_mem = nullptr;
    PoolView pool{};
    int* d_compact = nullptr;  // drain-phase compaction lists (kernels/wavefront.cu k_compact_*)
    // scratch pool of the parity entry points (FAST mode runs the production kernels over it)
    int scratch_cap = 0;
    void* scratch_mem = nullptr;
    PoolView scratch{};
    size_t smem_limit = 0;  // opt-in shared memory per block
    Control* d_ctl = nullptr;
    Control* h_ctl = nullptr;  // pinned, 2 entries

    // per-kernel timing (trt_opts.time_kernels)
    std::vector<cudaEvent_t> marks;  // 6 per timed iteration
    size_t marks_used = 0;
    int marks_mask = 0x3f;
    trt_kernel_times ktimes{};

    // scratch accumulation buffer for trt_render_to_host
    float* d_accum_own = nullptr;
    size_t accum_own_bytes = 0;
};

namespace {

float __int_as_float_host(int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

int use_device(trt_ctx* c) {
    CU(cudaSetDevice(c->device));
    return 0;
}

void free_scene(trt_ctx* c) {
    for (auto t : c->tex_objs) cudaDestroyTextureObject(t);
    for (auto a : c->tex_arrays) cudaFreeArray(a);
    c->tex_objs.clear();
    c->tex_arrays.clear();
    cudaFree(c->d_objects);
    cudaFree(c->d_ref_nodes);
    cudaFree(c->d_lights);
    if (c->wide_from_pool) {
        cudaFreeAsync(c->d_wide_nodes, c->stream);
        cudaFreeAsync(c->d_tris, c->stream);
    } else {
        cudaFree(c->d_wide_nodes);
        cudaFree(c->d_tris);
    }
    c->wide_from_pool = false;
    cudaFree(c->d_cnodes);
    c->d_cnodes = nullptr;
    c->d_objects = c->d_ref_nodes = c->d_wide_nodes = c->d_tris = nullptr;
    c->d_lights = nullptr;
    c->have_scene = false;
}

// texture object with the reference's descriptors (reference src/renderer.cu:97-126)
int make_texture(trt_ctx* c, const trt_image& img) {
    const int w = img.width, h = img.height;
    std::vector<unsigned char> rgba((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        rgba[i * 4 + 0] = img.rgb[i * 3 + 0];
        rgba[i * 4 + 1] = img.rgb[i * 3 + 1];
        rgba[i * 4 + 2] = img.rgb[i * 3 + 2];
        rgba[i * 4 + 3] = 255;
    }
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr;
    CU(cudaMallocArray(&arr, &desc, w, h));
    c->tex_arrays.push_back(arr);
    CU(cudaMemcpy2DToArray(arr, 0, 0, rgba.data(), (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeWrap;
    td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    CU(cudaCreateTextureObject(&obj, &res, &td, nullptr));
    c->tex_objs.push_back(obj);
    return 0;
}

int ensure_rng_tables(trt_ctx* c, int w, int h) {
    if (c->rng_w == w && c->rng_h == h) return 0;
    cudaFree(c->d_row_a);
    cudaFree(c->d_row_b);
    cudaFree(c->d_col_a);
    cudaFree(c->d_col_b);
    c->d_row_a = c->d_row_b = c->d_col_a = c->d_col_b = nullptr;
    c->rng_w = c->rng_h = 0;
    std::vector<Gf2Mat> rows, col_lo, col_hi;
    xorwow_build_row_matrices(w, h, rows);
    xorwow_build_col_levels(w, col_lo, col_hi);
    std::vector<uint32_t> ta((size_t)h * kXwWindowEntries * 4), tb((size_t)h * kXwWindowEntries);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < h; r++)
        xorwow_window_table(rows[r], ta.data() + (size_t)r * kXwWindowEntries * 4, tb.data() + (size_t)r * kXwWindowEntries);
    CU(cudaMalloc(&c->d_row_a, ta.size() * 4));
    CU(cudaMalloc(&c->d_row_b, tb.size() * 4));
    CU(cudaMemcpyAsync(c->d_row_a, ta.data(), ta.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_row_b, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice, c->stream));
    const size_t n_col = col_lo.size() + col_hi.size();
    std::vector<uint32_t> ca(n_col * kXwWindowEntries * 4), cb(n_col * kXwWindowEntries);
    for (size_t t = 0; t < n_col; t++)
        xorwow_window_table(t < col_lo.size() ? col_lo[t] : col_hi[t - col_lo.size()], ca.data() + t * kXwWindowEntries * 4,
                            cb.data() + t * kXwWindowEntries);
    CU(cudaMalloc(&c->d_col_a, ca.size() * 4));
    CU(cudaMalloc(&c->d_col_b, cb.size() * 4));
    CU(cudaMemcpyAsync(c->d_col_a, ca.data(), ca.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_col_b, cb.data(), cb.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->rng_w = w;
    c->rng_h = h;
    return 0;
}

int ensure_col_vecs(trt_ctx* c, size_t entries) {
    if (entries <= c->col_vecs_cap) return 0;
    cudaFree(c->d_col_vecs);
    c->d_col_vecs = nullptr;
    c->col_vecs_cap = 0;
    CU(cudaMalloc(&c->d_col_vecs, entries * sizeof(XwColVec)));
    c->col_vecs_cap = entries;
    return 0;
}

// One allocation, carved into the arrays of PoolView (+ the dead mask of the render pool).
int alloc_pool(int cap, bool render_pool, void** mem, PoolView* pool) {
    const size_t n = (size_t)cap;
    // od (32) + rs (32) + thr, pend, sh_d (16 each) + hit (8) = 120 bytes per slot + dead mask (1 bit per slot)
    const size_t bytes = n * (2 * 32 + 3 * 16 + 8) + (render_pool ? n / 8 + 64 : 0);
    CU(cudaMalloc(mem, bytes));
    char* p = (char*)*mem;
    auto take = [&](size_t b) { void* r = p; p += b; return r; };
    pool->od = (float4*)take(n * 32);
    pool->rs = (uint4*)take(n * 32);
    pool->thr = (float4*)take(n * 16);
    pool->pend = (float4*)take(n * 16);
    pool->sh_d = (float4*)take(n * 16);
    pool->hit = (float2*)take(n * 8);
    pool->dead_mask = render_pool ? (uint32_t*)take(n / 8 + 64) : nullptr;
    pool->capacity = cap;
    return 0;
}

int ensure_pool(trt_ctx* c, int cap) {
    cap = std::max(512, (cap + 511) & ~511);
    if (c->pool_cap == cap) return 0;
    cudaFree(c->pool_mem);
    cudaFree(c->d_compact);
    c->pool_mem = nullptr;
    c->d_compact = nullptr;
    c->pool_cap = 0;
    if (int rc = alloc_pool(cap, true, &c->pool_mem, &c->pool)) return rc;
    // two lists of up to cap entries each (live slots beyond the new bound / dead slots below it; the bound is
    // clamped to kCompactMinCap, so with a small pool the dead list can hold nearly the whole pool)
    CU(cudaMalloc(&c->d_compact, 2 * ((size_t)cap + 512) * sizeof(int)));
    c->pool_cap = cap;
    return 0;
}

int ensure_scratch(trt_ctx* c, int n) {
    const int cap = std::max(256, (n + 255) & ~255);
    if (c->scratch_cap >= cap) {
        c->scratch.capacity = cap;
        return 0;
    }
    cudaFree(c->scratch_mem);
    c->scratch_mem = nullptr;
    c->scratch_cap = 0;
    if (int rc = alloc_pool(cap, false, &c->scratch_mem, &c->scratch)) return rc;
    c->scratch_cap = cap;
    return 0;
}

// Launch configuration of the persistent traversal kernels.  Defaults from the B200 sweeps
// (profiles/): 768-thread CTAs (24 warps/SM at 80 registers); whatever shared memory the stacks,
// staging buffers and ray queues leave goes to the top of the tree (padded node staging);
// phases of two node steps followed by triangle steps.
// TRT_FAST_THREADS / TRT_SMEM_NODES / TRT_REFILL / TRT_PHASES override (tuning).
LaunchDims launch_dims(const trt_ctx* c) {
    LaunchDims d;
    d.sms = c->sms;
    d.fast_threads = 896;  // 28 warps per SM at 70 registers: the traversal is issue bound and wants warps (512 -> 768 -> 896 threads: 59.3 / 50.4 / 48.4 ms per 32 spp)
    if (const char* e = getenv("TRT_FAST_THREADS")) {
        const int v = atoi(e);
        if (v == 512 || v == 768 || v == 896 || v == 1024) d.fast_threads = v;
    }
    const int fit = wf_fast_max_smem_nodes(d.fast_threads, c->smem_limit);
    d.smem_nodes = fit;
    if (const char* e = getenv("TRT_SMEM_NODES")) d.smem_nodes = std::max(0, std::min(fit, atoi(e)));
    // trees beyond a few MB: the node fetch is bound by L1 requests; the upload has compressed the nodes to 64 bytes
    // (two 256-bit loads per node step) and the kernels read that form, nothing staged
    d.wide_loads = c->sc.cnodes != nullptr;
    if (d.wide_loads) d.fast_threads = 896;  // the compressed-node kernels exist for 896-thread CTAs only
    d.refill_below = 32;
    d.shade_block = 128;
    d.shade_minb = 9;  // 56 registers, 8 bytes spilled: 1152 threads per SM; shade waits on memory (48.8 -> 46.9 ms per 64 spp; 10 CTAs = 48 registers spill 108 bytes: 48.6)
    if (const char* e = getenv("TRT_SHADE_MINB")) d.shade_minb = atoi(e);
    if (const char* e = getenv("TRT_SHADE_BLOCK")) { const int v = atoi(e); if (v == 64 || v == 128 || v == 256 || v == 512) d.shade_block = v; }
    d.merged_trace = true;
    if (const char* e = getenv("TRT_MERGED_TRACE")) d.merged_trace = atoi(e) != 0;
    d.finish_below = 128 << 10;
    if (const char* e = getenv("TRT_FINISH_BELOW")) d.finish_below = std::max(0, atoi(e));
    d.debug_checks = false;
    if (const char* e = getenv("TRT_DEBUG_CHECKS")) d.debug_checks = atoi(e) != 0;
    d.compact_quarters = 3;
    if (const char* e = getenv("TRT_COMPACT_QUARTERS")) d.compact_quarters = std::max(1, std::min(3, atoi(e)));
    if (const char* e = getenv("TRT_REFILL")) d.refill_below = std::max(1, std::min(32, atoi(e)));
    int ranked = 0;  // measured equal to the brute-force pass on C2 (65.4 vs 65.7 ms of extend per 64 spp): kept as an option
    if (const char* e = getenv("TRT_RANKED_TOP")) ranked = atoi(e) != 0;
    d.closest_phases = Phases{2, 12, 8, ranked};
    d.shadow_phases = Phases{2, 12, 8, 0};
    if (d.wide_loads) {  // deep trees (C5: 40 node steps per ray, 5 lanes per triangle step): longer node phases, -6 %
        d.closest_phases = Phases{4, 12, 8, ranked};
        d.shadow_phases = Phases{4, 12, 8, 0};
    }
    if (const char* e = getenv("TRT_PHASES")) {  // "iters,node_min,tri_min[,iters,node_min,tri_min]" closest[,shadow]
        int v[6] = {1, 33, 33, 1, 33, 33};
        const int n = sscanf(e, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
        if (n >= 3) d.closest_phases = d.shadow_phases = Phases{std::max(1, v[0]), v[1], std::max(1, v[2]), 0};
        if (n >= 6) d.shadow_phases = Phases{std::max(1, v[3]), v[4], std::max(1, v[5]), 0};
        d.closest_phases.ranked_top = ranked;
    }
    return d;
}

void fill_job(trt_ctx* c, JobParams& job, float* d_accum, int w, int h, int first_frame_seed, int n_frames,
              int stride, const void* cam, const trt_opts& o) {
    memcpy(&job.cam, cam, sizeof(Camera));
    job.rc.width = w;
    job.rc.height = h;
    job.rc.max_depth = o.max_depth;
    job.rc.rr_threshold = o.rr_threshold;
    job.first_frame_seed = first_frame_seed;
    job.frame_stride = stride;
    job.seed_base = o.seed_base;
    job.n_frames = n_frames;
    job.row_a = reinterpret_cast<const uint4*>(c->d_row_a);
    job.row_b = c->d_row_b;
    job.col_vecs = c->d_col_vecs;
    job.accum = d_accum;
}

// Sorts the root-level list by the axis on which each leaf box is thinnest and fills the per-axis
// counts and planes the shadow top phase reads (kernels/common.cuh TopPrims).
void finalize_top(TopPrims& tp) {
    struct Item { float4 v0, e1, e2, bmin, bmax; int axis; int small; };
    std::vector<Item> items(tp.n);
    // small primitives (lights): box area below 2 % of the box around the whole list and the tree; at most 4
    float lo[3] = {tp.root_lo.x, tp.root_lo.y, tp.root_lo.z}, hi[3] = {tp.root_hi.x, tp.root_hi.y, tp.root_hi.z};
    auto area = [](const float* a, const float* b) {
        const float x = std::max(b[0] - a[0], 0.f), y = std::max(b[1] - a[1], 0.f), z = std::max(b[2] - a[2], 0.f);
        return 2.f * (x * y + y * z + z * x);
    };
    for (int i = 0; i < tp.n; i++) {
        const float bl[3] = {tp.bmin[i].x, tp.bmin[i].y, tp.bmin[i].z}, bh[3] = {tp.bmax[i].x, tp.bmax[i].y, tp.bmax[i].z};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], bl[k]); hi[k] = std::max(hi[k], bh[k]); }
    }
    const float scene_area = area(lo, hi);
    int n_small = 0;
    for (int i = 0; i < tp.n; i++) {
        const float ext[3] = {tp.bmax[i].x - tp.bmin[i].x, tp.bmax[i].y - tp.bmin[i].y, tp.bmax[i].z - tp.bmin[i].z};
        int ax = 0;
        for (int k = 1; k < 3; k++)
            if (ext[k] < ext[ax]) ax = k;
        const float bl[3] = {tp.bmin[i].x, tp.bmin[i].y, tp.bmin[i].z}, bh[3] = {tp.bmax[i].x, tp.bmax[i].y, tp.bmax[i].z};
        const int small = (area(bl, bh) < 0.02f * scene_area && n_small < 4) ? 1 : 0;
        n_small += small;
        items[i] = Item{tp.v0[i], tp.e1[i], tp.e2[i], tp.bmin[i], tp.bmax[i], ax, small};
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
        return a.axis != b.axis ? a.axis < b.axis : a.small < b.small;
    });
    tp.n_axis[0] = tp.n_axis[1] = tp.n_axis[2] = 0;
    tp.n_thin[0] = tp.n_thin[1] = tp.n_thin[2] = 0;
    tp.n_full = 0;
    for (int i = 0; i < tp.n; i++) {
        const Item& it = items[i];
        tp.v0[i] = it.v0; tp.e1[i] = it.e1; tp.e2[i] = it.e2; tp.bmin[i] = it.bmin; tp.bmax[i] = it.bmax;
        tp.thin_lo[i] = it.axis == 0 ? it.bmin.x : (it.axis == 1 ? it.bmin.y : it.bmin.z);
        tp.thin_hi[i] = it.axis == 0 ? it.bmax.x : (it.axis == 1 ? it.bmax.y : it.bmax.z);
        tp.thin2[i] = make_float2(tp.thin_lo[i], tp.thin_hi[i]);
        tp.n_axis[it.axis]++;
        if (it.small) tp.full_idx[tp.n_full++] = i;
        else tp.n_thin[it.axis]++;
    }
    // pairs in ascending object index (kernels/common.cuh TopPrims::pair_*)
    std::vector<int> order(tp.n);
    for (int i = 0; i < tp.n; i++) order[i] = i;
    auto obj_id = [&](int i) { int v; memcpy(&v, &tp.v0[i].w, 4); return v & 0x3fffffff; };
    std::sort(order.begin(), order.end(), [&](int a, int b) { return obj_id(a) < obj_id(b); });
    tp.n_pairs = (tp.n + 1) / 2;
    for (int j = 0; j < tp.n_pairs; j++) {
        float v[2][3] = {{0, 0, 0}, {0, 0, 0}}, a[2][3] = {{0, 0, 0}, {0, 0, 0}}, b[2][3] = {{0, 0, 0}, {0, 0, 0}};
        int id[2] = {-1, -1};
        for (int k = 0; k < 2 && 2 * j + k < tp.n; k++) {
            const int i = order[2 * j + k];
            v[k][0] = tp.v0[i].x; v[k][1] = tp.v0[i].y; v[k][2] = tp.v0[i].z;
            a[k][0] = tp.e1[i].x; a[k][1] = tp.e1[i].y; a[k][2] = tp.e1[i].z;
            b[k][0] = tp.e2[i].x; b[k][1] = tp.e2[i].y; b[k][2] = tp.e2[i].z;
            memcpy(&id[k], &tp.v0[i].w, 4);
        }
        for (int c = 0; c < 3; c++) {
            tp.pair_nv0[j][c] = make_float2(-v[0][c], -v[1][c]);
            tp.pair_e1[j][c] = make_float2(a[0][c], a[1][c]);
            tp.pair_ne1[j][c] = make_float2(-a[0][c], -a[1][c]);
            tp.pair_e2[j][c] = make_float2(b[0][c], b[1][c]);
        }
        tp.pair_id[j] = make_int2(id[0], id[1]);
    }
}

int check_opts(const trt_opts* in, trt_opts* o) {
    if (in) *o = *in;
    else trt_default_opts(o);
    if (o->max_depth < 1 || o->max_depth > 255) return fail(TRT_ERR_ARG, "max_depth %d out of range [1,255]", o->max_depth);
    if (o->traversal != TRT_TRAVERSE_FAST && o->traversal != TRT_TRAVERSE_REF)
        return fail(TRT_ERR_ARG, "unknown traversal mode %d", o->traversal);
    if (o->pool_paths < 0) return fail(TRT_ERR_ARG, "pool_paths must be >= 0");
    return 0;
}

int render_impl(trt_ctx* c, float* d_accum, int w, int h, int first, int n_frames, int stride, const void* cam,
                const trt_opts* opts_in) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "trt_render before trt_upload_scene");
    if (!d_accum || !cam) return fail(TRT_ERR_ARG, "null accum / camera pointer");
    if (w <= 0 || h <= 0 || n_frames < 0 || stride < 1) return fail(TRT_ERR_ARG, "bad render dimensions");
    if ((long long)w * h > (1ll << 30)) return fail(TRT_ERR_ARG, "image too large");
    trt_opts o;
    if (int rc = check_opts(opts_in, &o)) return rc;
    if (o.traversal == TRT_TRAVERSE_REF && c->sc.n_ref_nodes == 0)
        return fail(TRT_ERR_STATE, "TRT_TRAVERSE_REF needs the reference node array (scene was built on the device without it)");
    if ((long long)o.seed_base + first < 0) return fail(TRT_ERR_ARG, "negative RNG seed");
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_rng_tables(c, w, h)) return rc;
    int kFrameChunk = kFrameChunkDefault;
    if (const char* e = getenv("TRT_FRAME_CHUNK")) kFrameChunk = std::max(1, std::min(1024, atoi(e)));
    // pool_paths = 0: size the pool to the job -- the largest power of two that the samples of one job (at most
    // kFrameChunk frames) fill, between 256 Ki and 32 Mi slots.  Every iteration has a fixed cost (launch gaps, the
    // start-up and tail of the persistent kernel), so fewer, larger iterations win: a job that fits the pool is one
    // refill and a drain.  B200, C2: 1 spp (2.1 M samples) 8.34 / 6.06 / 4.93 / 4.51 / 4.72 ms at 256 Ki / 512 Ki /
    // 1 Mi / 2 Mi / 4 Mi; 4 spp 3.41 / 3.19 / 3.07 ms/spp at 2 / 4 / 8 Mi; 16 spp 2.71 / 2.64 / 2.65 at 8 / 16 / 32 Mi;
    // 64 spp (133 M samples) 2.59 / 2.50 / 2.47 at 8 / 16 / 32 Mi; C1 (4.9 M samples) 0.331 / 0.302 / 0.299 / 0.309 at
    // 1 / 2 / 4 / 8 Mi
    int pool_paths = o.pool_paths;
    if (pool_paths == 0) {
        const unsigned long long job = (unsigned long long)w * h * (unsigned long long)std::min(n_frames, kFrameChunk);
        pool_paths = 256 << 10;
        while (pool_paths < (32 << 20) && 2ull * (unsigned long long)pool_paths * 100 <= job * 105) pool_paths <<= 1;
    }
    if (int rc = ensure_pool(c, pool_paths)) return rc;
    if (int rc = ensure_col_vecs(c, (size_t)std::min(n_frames, kFrameChunk) * w)) return rc;

    c->marks_used = 0;
    IterStreams st;
    st.main = c->stream;
    st.mark_mask = o.time_kernels == 2 ? 0x0c : 0x3f;  // 2: only the marks around the extend kernel
    c->marks_mask = st.mark_mask;
    bool compact = true;
    if (const char* e = getenv("TRT_COMPACT")) compact = atoi(e) != 0;
    const LaunchDims dims = launch_dims(c);
    CU(cudaEventRecord(c->ev_begin, c->stream));
    const unsigned long long pixels = (unsigned long long)w * h;
    for (int f0 = 0; f0 < n_frames; f0 += kFrameChunk) {
        const int nf = std::min(kFrameChunk, n_frames - f0);
        JobParams job;
        fill_job(c, job, d_accum, w, h, first + f0 * stride, nf, stride, cam, o);
        wf_col_table(reinterpret_cast<const uint4*>(c->d_col_a), c->d_col_b, w, job.first_frame_seed, stride, o.seed_base, nf, c->d_col_vecs,
                     c->stream);
        wf_init_pool(c->pool, c->stream);
        wf_begin_job(c->d_ctl, pixels * nf, c->pool_cap, c->stream);
        c->launches += 3;
        st.visit_cap = c->pool_cap;
        st.samples_left = true;
        st.mostly_live = true;
        st.finish_below = 0;
        // Issue batches of iterations, staying one batch ahead of the completion poll.
        // near_drain: the job can reach its drain phase within the batches in flight (the poll is up to two
        // batches old; an iteration starts about capacity / 6 samples) -- from here on the batches are short.
        // compact_near: the same with the short batches (poll at most 8 iterations old) -- from here on the
        // compaction kernels ride along; late by an iteration or two only costs time, never correctness
        bool near_drain = pixels * nf <= 8ull * (unsigned long long)c->pool_cap;
        bool compact_near = pixels * nf <= 2ull * (unsigned long long)c->pool_cap;
        auto issue = [&](int slot, int n_iter) -> int {
            for (int i = 0; i < n_iter; i++) {
                cudaEvent_t* marks = nullptr;
                if (o.time_kernels) {
                    while (c->marks.size() < c->marks_used + 6) {
                        cudaEvent_t e;
                        CU(cudaEventCreate(&e));
                        c->marks.push_back(e);
                    }
                    marks = c->marks.data() + c->marks_used;
                    c->marks_used += 6;
                }
                c->launches += (unsigned long long)wf_iteration(c->pool, c->d_ctl, c->sc, c->top, job, o.traversal,
                                                                o.count_rays != 0, dims, st, marks,
                                                                compact_near && compact ? c->d_compact : nullptr);
            }
            CU(cudaMemcpyAsync(&c->h_ctl[slot], c->d_ctl, sizeof(Control), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaEventRecord(c->ev_poll[slot], c->stream));
            return 0;
        };
        // every iteration advances each live path by one vertex, so the job needs at most
        // (samples / pool + 1) * (max_depth + 2) iterations; anything beyond that is a bug
        const unsigned long long max_iterations =
            ((pixels * nf) / (unsigned long long)c->pool_cap + 2) * (unsigned long long)(o.max_depth + 2) +
            8 * kBatchIterations;
        unsigned long long issued = 0, b = 0;
        int batch = near_drain ? kTailBatchIterations : kBatchIterations;
        if (int rc = issue(0, batch)) return rc;
        issued += (unsigned long long)batch;
        for (;;) {
            if (int rc = issue((int)((b + 1) & 1), batch)) return rc;
            issued += (unsigned long long)batch;
            CU(cudaEventSynchronize(c->ev_poll[b & 1]));
            const Control& hc = c->h_ctl[b & 1];
            if (hc.alive == 0 && hc.next_sample == hc.total_samples) break;
            if (hc.total_samples - hc.next_sample <= 8ull * (unsigned long long)c->pool_cap) near_drain = true;
            if (hc.total_samples - hc.next_sample <= 2ull * (unsigned long long)c->pool_cap) compact_near = true;
            // what the poll says about the rest of the job: the shade grid follows the compacted bound,
            // regeneration stops with the last sample; towards the drain phase the batches are short, so that
            // the bound is fresh and few iterations are queued behind the last live path
            st.visit_cap = std::min(st.visit_cap, hc.active_cap);
            if (hc.next_sample == hc.total_samples) st.samples_left = false;
            st.mostly_live = (long long)hc.alive * 2 > (long long)hc.active_cap;
            // the poll is up to two batches old and the live paths shrink by a quarter per iteration: let the tail
            // kernel ride along from well above its threshold (it decides on the device)
            st.finish_below = (!st.samples_left && (long long)hc.alive <= 16ll * dims.finish_below) ? dims.finish_below : 0;
            if (near_drain) batch = kTailBatchIterations;
            // once the tail kernel rides along the job can end in any iteration: single-iteration batches, so that at
            // most two empty iterations are queued behind it (45 us each; a 1-spp call is 4.4 ms)
            if (st.finish_below > 0) batch = 1;
            b++;
            if (issued > max_iterations) {
                cudaStreamSynchronize(c->stream);
                return fail(TRT_ERR_STATE, "wavefront did not drain after %llu iterations (alive=%d, next=%llu of %llu)",
                            issued, hc.alive, hc.next_sample, hc.total_samples);
            }
        }
    }
    CU(cudaEventRecord(c->ev_end, c->stream));
    CU(cudaGetLastError());
    return 0;
}

int upload_core(trt_ctx* c, const Object* objs, float4* d_objs_ready, int n_objects, const LinearBVHNode* nd, int n_nodes,
                bool have_ref, const int* lights, int n_lights, const trt_image* textures, int n_textures, int builder);
}  // namespace

extern "C" {

const char* trt_last_error(void) { return g_err.c_str(); }
const char* trt_version(void) { return "tryraytrace_b200 0.1 (sm_100a)"; }

void trt_default_opts(trt_opts* o) {
    memset(o, 0, sizeof(*o));
    o->max_depth = 30;
    o->rr_threshold = 3;
    o->seed_base = 1984;
    o->traversal = TRT_TRAVERSE_FAST;
    o->pool_paths = 0;
    o->count_rays = 0;
}

namespace {
int init_ctx(trt_ctx* c, int device) {
    c->device = device;
    CU(cudaSetDevice(device));
    CU(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    int smem_optin = 0;
    CU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    c->smem_limit = (size_t)smem_optin;
    if (wf_configure() != 0) {
        cudaGetLastError();
        return fail(TRT_ERR_CUDA, "cannot opt in to %d bytes of shared memory per block", smem_optin);
    }
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaEventCreate(&c->ev_begin));
    CU(cudaEventCreate(&c->ev_end));
    CU(cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming));
    CU(cudaMalloc(&c->d_ctl, sizeof(Control)));
    CU(cudaMemset(c->d_ctl, 0, sizeof(Control)));
    CU(cudaMallocHost(&c->h_ctl, 2 * sizeof(Control)));
    return 0;
}
}  // namespace

int trt_create(int device, trt_ctx** out) {
    if (!out) return fail(TRT_ERR_ARG, "null out pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(TRT_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= n) return fail(TRT_ERR_ARG, "device %d out of range (have %d)", device, n);
    trt_ctx* c = new trt_ctx();
    if (int rc = init_ctx(c, device)) {  // a half-built context is torn down again (message kept)
        const std::string keep = g_err;
        trt_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return 0;
}

int trt_destroy(trt_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_scene(c);
    cudaFree(c->d_row_a);
    cudaFree(c->d_row_b);
    cudaFree(c->d_col_a);
    cudaFree(c->d_col_b);
    cudaFree(c->d_col_vecs);
    cudaFree(c->pool_mem);
    cudaFree(c->d_compact);
    cudaFree(c->scratch_mem);
    cudaFree(c->d_ctl);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    cudaFree(c->d_accum_own);
    for (cudaEvent_t e : c->marks) cudaEventDestroy(e);
    for (cudaEvent_t e : {c->ev_begin, c->ev_end, c->ev_poll[0], c->ev_poll[1]})
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    cudaGetLastError();
    delete c;
    return 0;
}

int trt_upload_scene_ex(trt_ctx* c, const void* objects, int n_objects, const void* nodes, int n_nodes,
                        const int* lights, int n_lights, const trt_image* textures, int n_textures, int builder) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (!objects || n_objects <= 0) return fail(TRT_ERR_ARG, "empty object array");
    if (builder != TRT_BUILD_AUTO && builder != TRT_BUILD_HOST_SAH && builder != TRT_BUILD_DEVICE_LBVH)
        return fail(TRT_ERR_ARG, "unknown builder %d", builder);
    if (const char* e = getenv("TRT_BUILDER")) {  // tuning override: 1 = host SAH, 2 = device LBVH
        const int v = atoi(e);
        if (v == TRT_BUILD_HOST_SAH || v == TRT_BUILD_DEVICE_LBVH) builder = v;
    }
    if (builder == TRT_BUILD_AUTO)
        builder = (!nodes || n_objects > kAutoDeviceBuildAbove) ? TRT_BUILD_DEVICE_LBVH : TRT_BUILD_HOST_SAH;
    const bool have_ref = nodes && n_nodes > 0;
    if (!have_ref && builder != TRT_BUILD_DEVICE_LBVH)
        return fail(TRT_ERR_ARG, "empty node array (only TRT_BUILD_DEVICE_LBVH can build without the reference nodes)");
    if (n_lights < 0 || (n_lights > 0 && !lights)) return fail(TRT_ERR_ARG, "bad light list");
    if (n_textures < 0 || n_textures > 5) return fail(TRT_ERR_ARG, "at most 5 textures (MAX_TEXTURES)");
    if (n_objects >= (1 << 29)) return fail(TRT_ERR_ARG, "too many objects");
    const Object* objs = (const Object*)objects;
    const LinearBVHNode* nd = (const LinearBVHNode*)nodes;
    // validate what the kernels will index with
    for (int i = 0; i < n_lights; i++)
        if (lights[i] < 0 || lights[i] >= n_objects) return fail(TRT_ERR_ARG, "light index %d out of range", lights[i]);
    for (int i = 0; i < n_objects; i++)
        if (objs[i].tex_id >= n_textures) return fail(TRT_ERR_ARG, "object %d uses texture %d, only %d given", i, objs[i].tex_id, n_textures);
    for (int i = 0; have_ref && i < n_nodes; i++) {
        const LinearBVHNode& n = nd[i];
        if (n.is_leaf) {
            if (n.primitive_offset < 0 || n.primitive_count < 0 || n.primitive_offset + n.primitive_count > n_objects)
                return fail(TRT_ERR_ARG, "node %d: primitive range out of bounds", i);
        } else if (n.left_child_idx <= i || n.right_child_idx <= i || n.left_child_idx >= n_nodes ||
                   n.right_child_idx >= n_nodes) {
            return fail(TRT_ERR_ARG, "node %d: child index out of order", i);
        }
    }
    if (int rc = use_device(c)) return rc;
    return upload_core(c, objs, nullptr, n_objects, nd, n_nodes, have_ref, lights, n_lights, textures, n_textures, builder);
}

}  // extern "C"

namespace {
// Everything behind the argument checks.  `objs` is the host object array, or nullptr when `d_objs_ready` already
// holds the objects on the device (trt_upload_instanced: ownership passes to the context; device builder only).
int upload_core(trt_ctx* c, const Object* objs, float4* d_objs_ready, int n_objects, const LinearBVHNode* nd, int n_nodes,
                bool have_ref, const int* lights, int n_lights, const trt_image* textures, int n_textures, int builder) {
    CU(cudaStreamSynchronize(c->stream));
    free_scene(c);

    if (d_objs_ready) {
        c->d_objects = d_objs_ready;
    } else {
        CU(cudaMalloc(&c->d_objects, (size_t)n_objects * sizeof(Object)));
        CU(cudaMemcpy(c->d_objects, objs, (size_t)n_objects * sizeof(Object), cudaMemcpyHostToDevice));
    }
    if (have_ref) {
        CU(cudaMalloc(&c->d_ref_nodes, (size_t)n_nodes * sizeof(LinearBVHNode)));
        CU(cudaMemcpy(c->d_ref_nodes, nd, (size_t)n_nodes * sizeof(LinearBVHNode), cudaMemcpyHostToDevice));
    }
    if (n_lights) {
        CU(cudaMalloc(&c->d_lights, (size_t)n_lights * sizeof(int)));
        CU(cudaMemcpy(c->d_lights, lights, (size_t)n_lights * sizeof(int), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < n_textures; i++) {
        if (!textures[i].rgb || textures[i].width <= 0 || textures[i].height <= 0)
            return fail(TRT_ERR_ARG, "texture %d is empty", i);
        if (int rc = make_texture(c, textures[i])) return rc;
    }

    // re-layout for the fast path: wide BVH over the reference leaf boxes + triangle records
    TopPrims& tp = c->top;
    memset(&tp, 0, sizeof(tp));
    int n_wide = 0, n_tris = 0, n_top = 0, depth = 0, n_underivable = 0;
    float build_ms = 0.f;
    if (builder == TRT_BUILD_DEVICE_LBVH) {
        int max_leaf = 4;
        if (const char* e = getenv("TRT_MAX_LEAF")) max_leaf = std::max(1, std::min(4, atoi(e)));
        DeviceWideBvh dw;
        std::string err;
        bool top_sah = true;
        if (const char* e = getenv("TRT_TOP_SAH")) top_sah = atoi(e) != 0;
        if (build_wide_bvh_device(c->d_objects, n_objects, c->d_ref_nodes, have_ref ? n_nodes : 0, max_leaf, top_sah, &dw,
                                  c->stream, &err) != 0) {
            cudaFreeAsync(dw.d_nodes, c->stream);
            cudaFreeAsync(dw.d_tris, c->stream);
            cudaGetLastError();
            return fail(TRT_ERR_CUDA, "device BVH build: %s", err.c_str());
        }
        c->d_wide_nodes = dw.d_nodes;
        c->d_tris = dw.d_tris;
        c->wide_from_pool = true;
        tp = dw.top;
        n_wide = dw.n_nodes; n_tris = dw.n_tris; n_top = dw.n_top; depth = dw.depth; n_underivable = dw.n_underivable;
        build_ms = dw.build_ms;
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        WideBvh wb;
        build_wide_bvh(objs, n_objects, nd, n_nodes, wb);
        build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        CU(cudaMalloc(&c->d_wide_nodes, std::max<size_t>(wb.nodes.size(), 1) * sizeof(WideNode)));
        CU(cudaMemcpy(c->d_wide_nodes, wb.nodes.data(), wb.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_tris, std::max<size_t>(wb.tris.size(), 1) * sizeof(TriRecord)));
        CU(cudaMemcpy(c->d_tris, wb.tris.data(), wb.tris.size() * sizeof(TriRecord), cudaMemcpyHostToDevice));
        tp.n = (int)wb.top.size();
        for (int i = 0; i < tp.n; i++) {
            const TopPrim& t = wb.top[i];
            tp.v0[i] = make_float4(t.v0[0], t.v0[1], t.v0[2], __int_as_float_host(t.id));
            tp.e1[i] = make_float4(t.e1[0], t.e1[1], t.e1[2], 0.f);
            tp.e2[i] = make_float4(t.e2[0], t.e2[1], t.e2[2], 0.f);
            tp.bmin[i] = make_float4(t.mn[0], t.mn[1], t.mn[2], 0.f);
            tp.bmax[i] = make_float4(t.mx[0], t.mx[1], t.mx[2], 0.f);
        }
        tp.root_lo = make_float4(wb.root_mn[0], wb.root_mn[1], wb.root_mn[2], 0.f);
        tp.root_hi = make_float4(wb.root_mx[0], wb.root_mx[1], wb.root_mx[2], 0.f);
        n_wide = (int)wb.nodes.size(); n_tris = (int)wb.tris.size(); n_top = wb.n_top_prims; depth = wb.depth;
        n_underivable = wb.n_underivable;
    }
    // Compressed nodes for trees far larger than the caches (north-star subsystem 1; kernels/traverse_fast.cuh CNode):
    // conservative 8-bit child boxes, 64 bytes per node.  TRT_COMPRESSED=1/0 forces / forbids it (tests, tuning).
    bool compress = (size_t)n_wide * sizeof(WideNode) > ((size_t)16 << 20);
    if (const char* e = getenv("TRT_COMPRESSED")) compress = atoi(e) != 0;
    if (compress && n_wide > 0) {
        int* d_bad = nullptr;
        int bad = 0;
        CU(cudaMalloc(&c->d_cnodes, (size_t)n_wide * 64));
        CU(cudaMalloc(&d_bad, sizeof(int)));
        CU(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
        wf_compress_nodes(c->d_wide_nodes, n_wide, c->d_cnodes, d_bad, c->stream);
        CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(d_bad);
        if (bad) {  // a box the grid cannot contain (not seen so far): keep the uncompressed nodes
            cudaFree(c->d_cnodes);
            c->d_cnodes = nullptr;
        } else {    // the kernels read only the compressed form from here on
            if (c->wide_from_pool) cudaFreeAsync(c->d_wide_nodes, c->stream);
            else cudaFree(c->d_wide_nodes);
            c->d_wide_nodes = nullptr;
        }
    }
    finalize_top(tp);
    if (3 * depth + 1 > kWideStackEntries) {
        free_scene(c);
        return fail(TRT_ERR_ARG, "wide BVH depth %d exceeds the traversal stack", depth);
    }

    SceneDev& sc = c->sc;
    memset(&sc, 0, sizeof(sc));
    sc.objects = c->d_objects;
    sc.ref_nodes = c->d_ref_nodes;
    sc.lights = c->d_lights;
    sc.n_objects = n_objects;
    sc.n_ref_nodes = have_ref ? n_nodes : 0;
    sc.n_lights = n_lights;
    sc.n_textures = n_textures;
    for (int i = 0; i < n_textures; i++) sc.tex[i] = c->tex_objs[i];
    sc.wide_nodes = c->d_wide_nodes;
    sc.cnodes = c->d_cnodes;
    sc.tris = c->d_tris;
    sc.n_wide_nodes = n_wide;
    sc.n_tris = n_tris;

    trt_scene_info& in = c->info;
    memset(&in, 0, sizeof(in));
    in.n_objects = n_objects;
    in.n_ref_nodes = sc.n_ref_nodes;
    in.n_lights = n_lights;
    in.n_textures = n_textures;
    in.n_wide_nodes = n_wide;
    in.n_wide_leaf_tris = n_tris;
    in.n_top_prims = n_top;
    in.wide_node_bytes = c->d_cnodes ? 64 : (int)sizeof(WideNode);
    in.tri_record_bytes = (int)sizeof(TriRecord);
    in.wide_depth = depth;
    in.builder = builder;
    in.build_ms = build_ms;
    in.n_underivable = n_underivable;
    c->have_scene = true;
    return 0;
}
}  // namespace

extern "C" {

int trt_upload_instanced(trt_ctx* c, const void* extra, int n_extra, const void* unit, int n_unit,
                         const float* instances, int n_instances, const int* lights, int n_lights,
                         const trt_image* textures, int n_textures) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (n_extra < 0 || n_unit <= 0 || n_instances <= 0 || !unit || !instances || (n_extra > 0 && !extra))
        return fail(TRT_ERR_ARG, "empty mesh / instance list");
    const long long total = (long long)n_extra + (long long)n_unit * n_instances;
    if (total >= (1ll << 29)) return fail(TRT_ERR_ARG, "too many objects");
    if (n_lights < 0 || (n_lights > 0 && !lights)) return fail(TRT_ERR_ARG, "bad light list");
    if (n_textures < 0 || n_textures > 5) return fail(TRT_ERR_ARG, "at most 5 textures (MAX_TEXTURES)");
    for (int i = 0; i < n_lights; i++)
        if (lights[i] < 0 || lights[i] >= total) return fail(TRT_ERR_ARG, "light index %d out of range", lights[i]);
    const Object* ex = (const Object*)extra;
    const Object* un = (const Object*)unit;
    for (int i = 0; i < n_extra; i++)
        if (ex[i].tex_id >= n_textures) return fail(TRT_ERR_ARG, "object %d uses texture %d, only %d given", i, ex[i].tex_id, n_textures);
    for (int i = 0; i < n_unit; i++)
        if (un[i].tex_id >= n_textures) return fail(TRT_ERR_ARG, "mesh triangle %d uses texture %d, only %d given", i, un[i].tex_id, n_textures);
    if (int rc = use_device(c)) return rc;
    // the mesh once, the placements, and a kernel that writes the 112-byte records where the reference's loader
    // would have put them (v' = fma(v, scale, offset), src/loader.cpp:51 as compiled with contraction)
    float4 *d_objs = nullptr, *d_unit = nullptr, *d_inst = nullptr;
    auto cleanup = [&] { cudaFree(d_unit); cudaFree(d_inst); };
    CU(cudaMalloc(&d_objs, (size_t)total * sizeof(Object)));
    if (cudaMalloc(&d_unit, (size_t)n_unit * sizeof(Object)) != cudaSuccess || cudaMalloc(&d_inst, (size_t)n_instances * 16) != cudaSuccess) {
        cleanup();
        cudaFree(d_objs);
        return fail(TRT_ERR_CUDA, "out of device memory for the instanced scene");
    }
    cudaMemcpyAsync(d_unit, unit, (size_t)n_unit * sizeof(Object), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_inst, instances, (size_t)n_instances * 16, cudaMemcpyHostToDevice, c->stream);
    if (n_extra) cudaMemcpyAsync(d_objs, extra, (size_t)n_extra * sizeof(Object), cudaMemcpyHostToDevice, c->stream);
    wf_instance_objects(d_unit, n_unit, d_inst, n_instances, d_objs + (size_t)n_extra * 7, c->stream);
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) {
        cudaFree(d_objs);
        return fail(TRT_ERR_CUDA, "instancing kernel: %s", cudaGetErrorString(e));
    }
    return upload_core(c, nullptr, d_objs, (int)total, nullptr, 0, false, lights, n_lights, textures, n_textures,
                       TRT_BUILD_DEVICE_LBVH);
}

int trt_get_objects(trt_ctx* c, void* out, int cap) {
    if (!c || !out) return fail(TRT_ERR_ARG, "null pointer");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "no scene uploaded");
    if (cap < c->sc.n_objects) return fail(TRT_ERR_ARG, "buffer too small: %d objects", c->sc.n_objects);
    if (int rc = use_device(c)) return rc;
    CU(cudaMemcpy(out, c->d_objects, (size_t)c->sc.n_objects * sizeof(Object), cudaMemcpyDeviceToHost));
    return c->sc.n_objects;
}

int trt_upload_scene(trt_ctx* c, const void* objects, int n_objects, const void* nodes, int n_nodes,
                     const int* lights, int n_lights, const trt_image* textures, int n_textures) {
    if (!nodes || n_nodes <= 0) return fail(TRT_ERR_ARG, "empty node array");
    return trt_upload_scene_ex(c, objects, n_objects, nodes, n_nodes, lights, n_lights, textures, n_textures,
                               TRT_BUILD_AUTO);
}

int trt_scene_info_get(trt_ctx* c, trt_scene_info* out) {
    if (!c || !out) return fail(TRT_ERR_ARG, "null pointer");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "no scene uploaded");
    *out = c->info;
    return 0;
}

int trt_render(trt_ctx* c, float* d_accum, int w, int h, int first, int n_frames, int stride, const void* cam,
               const trt_opts* opts) {
    return render_impl(c, d_accum, w, h, first, n_frames, stride, cam, opts);
}

int trt_render_to_host(trt_ctx* c, float* h_accum, int w, int h, int first, int n_frames, int stride,
                       const void* cam, const trt_opts* opts) {
    if (!c || !h_accum) return fail(TRT_ERR_ARG, "null pointer");
    if (w <= 0 || h <= 0) return fail(TRT_ERR_ARG, "bad render dimensions");
    if (int rc = use_device(c)) return rc;
    const size_t bytes = (size_t)w * h * 16;
    if (bytes > c->accum_own_bytes) {
        cudaFree(c->d_accum_own);
        c->d_accum_own = nullptr;
        c->accum_own_bytes = 0;
        CU(cudaMalloc(&c->d_accum_own, bytes));
        c->accum_own_bytes = bytes;
    }
    CU(cudaMemsetAsync(c->d_accum_own, 0, bytes, c->stream));
    if (int rc = render_impl(c, c->d_accum_own, w, h, first, n_frames, stride, cam, opts)) return rc;
    CU(cudaMemcpyAsync(h_accum, c->d_accum_own, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int trt_trace_primary(trt_ctx* c, int w, int h, int frame_seed, const void* cam, int traversal, int seed_base,
                      int* d_id, float* d_t, float* d_ray, uint32_t* d_fetched, uint32_t* d_entered,
                      uint32_t* d_tris) {
    if (!c || !cam) return fail(TRT_ERR_ARG, "null pointer");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "trt_trace_primary before trt_upload_scene");
    if (w <= 0 || h <= 0) return fail(TRT_ERR_ARG, "bad dimensions");
    if (traversal != TRT_TRAVERSE_FAST && traversal != TRT_TRAVERSE_REF) return fail(TRT_ERR_ARG, "unknown traversal mode");
    if (traversal == TRT_TRAVERSE_REF && c->sc.n_ref_nodes == 0) return fail(TRT_ERR_STATE, "TRT_TRAVERSE_REF needs the reference node array");
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_rng_tables(c, w, h)) return rc;
    if (int rc = ensure_col_vecs(c, (size_t)w)) return rc;
    trt_opts o;
    trt_default_opts(&o);
    o.seed_base = seed_base;
    JobParams job;
    fill_job(c, job, nullptr, w, h, frame_seed, 1, 1, cam, o);
    if (int rc = ensure_scratch(c, w * h)) return rc;
    wf_col_table(reinterpret_cast<const uint4*>(c->d_col_a), c->d_col_b, w, frame_seed, 1, seed_base, 1, c->d_col_vecs, c->stream);
    wf_trace_primary(c->sc, job, frame_seed, traversal, d_id, d_t, d_ray, d_fetched, d_entered, d_tris, c->top,
                     c->scratch, c->d_ctl, launch_dims(c), c->stream);
    c->launches += 2;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

int trt_trace_closest(trt_ctx* c, const float* d_rays, int n, int traversal, int* d_id, float* d_t) {
    if (!c || !d_rays || !d_id) return fail(TRT_ERR_ARG, "null pointer");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "no scene uploaded");
    if (n <= 0) return 0;
    if (traversal == TRT_TRAVERSE_REF && c->sc.n_ref_nodes == 0) return fail(TRT_ERR_STATE, "TRT_TRAVERSE_REF needs the reference node array");
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_scratch(c, n)) return rc;
    wf_trace_closest(c->sc, d_rays, n, traversal, d_id, d_t, c->top, c->scratch, c->d_ctl, launch_dims(c), c->stream);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

int trt_trace_shadow(trt_ctx* c, const float* d_rays, int n, int traversal, int* d_occ) {
    if (!c || !d_rays || !d_occ) return fail(TRT_ERR_ARG, "null pointer");
    if (!c->have_scene) return fail(TRT_ERR_STATE, "no scene uploaded");
    if (n <= 0) return 0;
    if (traversal == TRT_TRAVERSE_REF && c->sc.n_ref_nodes == 0) return fail(TRT_ERR_STATE, "TRT_TRAVERSE_REF needs the reference node array");
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_scratch(c, n)) return rc;
    wf_trace_shadow(c->sc, d_rays, n, traversal, d_occ, c->top, c->scratch, c->d_ctl, launch_dims(c), c->stream);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

int trt_rng_states(trt_ctx* c, int w, int h, int frame_seed, int seed_base, int first_pixel, int n,
                   uint32_t* d_states) {
    if (!c || !d_states) return fail(TRT_ERR_ARG, "null pointer");
    if (w <= 0 || h <= 0 || first_pixel < 0 || n < 0 || (long long)first_pixel + n > (long long)w * h)
        return fail(TRT_ERR_ARG, "pixel range out of bounds");
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_rng_tables(c, w, h)) return rc;
    if (int rc = ensure_col_vecs(c, (size_t)w)) return rc;
    trt_opts o;
    trt_default_opts(&o);
    o.seed_base = seed_base;
    JobParams job;
    Camera dummy;
    memset(&dummy, 0, sizeof(dummy));
    fill_job(c, job, nullptr, w, h, frame_seed, 1, 1, &dummy, o);
    wf_col_table(reinterpret_cast<const uint4*>(c->d_col_a), c->d_col_b, w, frame_seed, 1, seed_base, 1, c->d_col_vecs, c->stream);
    wf_rng_states(job, 0, first_pixel, n, d_states, c->stream);
    c->launches += 2;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

int trt_tonemap(trt_ctx* c, const float* d_accum, int w, int h, int frames, uint32_t* d_argb) {
    if (!c || !d_accum || !d_argb) return fail(TRT_ERR_ARG, "null pointer");
    if (w <= 0 || h <= 0 || frames <= 0) return fail(TRT_ERR_ARG, "bad dimensions");
    if (int rc = use_device(c)) return rc;
    wf_tonemap(d_accum, w * h, frames, d_argb, c->stream);
    c->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

int trt_tonemap_stream(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, void* cuda_stream) {
    if (!d_accum || !d_argb) return fail(TRT_ERR_ARG, "null pointer");
    if (n_pixels <= 0 || frames <= 0) return fail(TRT_ERR_ARG, "bad dimensions");
    wf_tonemap(d_accum, n_pixels, frames, d_argb, (cudaStream_t)cuda_stream);
    CU(cudaGetLastError());
    return 0;
}

int trt_synchronize(trt_ctx* c) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (int rc = use_device(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return 0;
}

int trt_get_counters(trt_ctx* c, trt_counters* out) {
    if (!c || !out) return fail(TRT_ERR_ARG, "null pointer");
    if (int rc = use_device(c)) return rc;
    Control hc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(&hc, c->d_ctl, sizeof(hc), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->samples = hc.cnt_samples;
    out->closest_rays = hc.cnt_closest;
    out->shadow_rays = hc.cnt_shadow;
    out->nodes_fetched = hc.cnt_nodes;
    out->tris_tested = hc.cnt_tris;
    out->replays = hc.cnt_replays;
    out->iterations = hc.cnt_iterations;
    out->kernel_launches = c->launches;
    out->nodes_closest = hc.cnt_nodes_closest;
    out->tris_closest = hc.cnt_tris_closest;
    out->tree_closest = hc.cnt_tree_closest;
    out->tree_shadow = hc.cnt_tree_shadow;
    out->check_violations = hc.cnt_violations;
    if (getenv("TRT_TRAV_STATS")) {  // lane-utilisation statistics of the traversal kernels (count_rays renders)
        const unsigned long long nodes_shadow = hc.cnt_nodes - hc.cnt_nodes_closest, tris_shadow = hc.cnt_tris - hc.cnt_tris_closest;
        auto line = [&](const char* name, const unsigned long long* d, unsigned long long rays, unsigned long long tree,
                        unsigned long long nodes, unsigned long long tris) {
            fprintf(stderr, "[trav] %s: rays %llu tree %llu (%.1f%%) chunks %llu rounds %llu lanes_with_ray/round %.2f | node steps %llu "
                    "(%.2f/round) lanes/node step %.2f | tri steps %llu (%.2f/round) lanes/tri step %.2f | nodes/tree ray %.2f tris/tree ray %.2f\n",
                    name, rays, tree, 100.0 * tree / (rays ? rays : 1), d[4], d[0], (double)d[1] / (d[0] ? d[0] : 1), d[2],
                    (double)d[2] / (d[0] ? d[0] : 1), (double)nodes / (d[2] ? d[2] : 1), d[3], (double)d[3] / (d[0] ? d[0] : 1),
                    (double)tris / (d[3] ? d[3] : 1), (double)nodes / (tree ? tree : 1), (double)tris / (tree ? tree : 1));
        };
        line("closest", hc.dbg, hc.cnt_closest, hc.cnt_tree_closest, hc.cnt_nodes_closest, hc.cnt_tris_closest);
        line("shadow ", hc.dbg + 8, hc.cnt_shadow, hc.cnt_tree_shadow, nodes_shadow, tris_shadow);
    }
    return 0;
}

int trt_reset_counters(trt_ctx* c) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (int rc = use_device(c)) return rc;
    wf_reset_counters(c->d_ctl, c->stream);
    CU(cudaStreamSynchronize(c->stream));
    c->launches = 0;
    return 0;
}

int trt_last_render_ms(trt_ctx* c, float* ms) {
    if (!c || !ms) return fail(TRT_ERR_ARG, "null pointer");
    if (int rc = use_device(c)) return rc;
    CU(cudaEventSynchronize(c->ev_end));
    CU(cudaEventElapsedTime(ms, c->ev_begin, c->ev_end));
    return 0;
}

int trt_kernel_times_get(trt_ctx* c, trt_kernel_times* out) {
    if (!c || !out) return fail(TRT_ERR_ARG, "null pointer");
    if (int rc = use_device(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    trt_kernel_times t;
    memset(&t, 0, sizeof(t));
    // TRT_ITER_LOG=<file>: one line per timed iteration (tuning aid; needs time_kernels = 1)
    FILE* log = nullptr;
    if (const char* e = getenv("TRT_ITER_LOG")) if (c->marks_mask == 0x3f) log = fopen(e, "w");
    for (size_t i = 0; i + 6 <= c->marks_used; i += 6) {  // marks: [0] prepare+regenerate [1], [2] extend [3] shade [4] shadow [5]
        float ms[4] = {0.f, 0.f, 0.f, 0.f};
        if (c->marks_mask == 0x3f) {
            CU(cudaEventElapsedTime(&ms[0], c->marks[i], c->marks[i + 1]));
            for (int k = 1; k < 4; k++) CU(cudaEventElapsedTime(&ms[k], c->marks[i + k + 1], c->marks[i + k + 2]));
        } else {
            CU(cudaEventElapsedTime(&ms[1], c->marks[i + 2], c->marks[i + 3]));
        }
        t.regen_ms += ms[0];  // k_refill
        t.extend_ms += ms[1];
        t.shade_ms += ms[2];
        t.shadow_ms += ms[3];
        t.iterations++;
        if (log) {
            float whole = 0.f;  // start of this iteration's extend to the start of the next one's
            if (i + 12 <= c->marks_used) cudaEventElapsedTime(&whole, c->marks[i + 2], c->marks[i + 8]);
            fprintf(log, "%d %.4f %.4f %.4f %.4f %.4f\n", t.iterations - 1, ms[0], ms[1], ms[2], ms[3], whole);
        }
    }
    if (log) fclose(log);
    *out = t;
    return 0;
}

void* trt_stream(trt_ctx* c) { return c ? (void*)c->stream : nullptr; }

int trt_set_stream(trt_ctx* c, void* cuda_stream) {
    if (!c) return fail(TRT_ERR_ARG, "null context");
    if (int rc = use_device(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return 0;
}

// ---- host surface ----------------------------------------------------------------------
int trt_load_obj(const char* filename, void* out, int cap, const float offset[3], float scale,
                 const float albedo[3], float metallic, float roughness) {
    if (!filename || !offset || !albedo) return fail(TRT_ERR_ARG, "null pointer");
    std::vector<Object> objs;
    load_obj(filename, objs, Vec{offset[0], offset[1], offset[2]}, scale, Vec{albedo[0], albedo[1], albedo[2]},
             metallic, roughness);
    if (out) {
        if ((int)objs.size() > cap) return fail(TRT_ERR_ARG, "buffer too small: %zu objects", objs.size());
        memcpy(out, objs.data(), objs.size() * sizeof(Object));
    }
    return (int)objs.size();
}

int trt_bvh_build(void* objects, int n, void* nodes, int cap) {
    if (!objects || !nodes || n <= 0) return fail(TRT_ERR_ARG, "bad arguments");
    if (cap < 2 * n - 1) return fail(TRT_ERR_ARG, "node buffer too small: need %d", 2 * n - 1);
    std::vector<Object> objs((Object*)objects, (Object*)objects + n);
    BVH bvh;
    bvh.build(objs);
    memcpy(objects, objs.data(), (size_t)n * sizeof(Object));
    memcpy(nodes, bvh.get_nodes().data(), bvh.get_nodes().size() * sizeof(LinearBVHNode));
    return (int)bvh.get_nodes().size();
}

int trt_collect_lights(const void* objects, int n, int* out, int cap) {
    if (!objects) return fail(TRT_ERR_ARG, "null pointer");
    const Object* o = (const Object*)objects;
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const Vec& e = o[i].emission;
        if (e.x > 0.1f || e.y > 0.1f || e.z > 0.1f) {
            if (out && cnt < cap) out[cnt] = i;
            cnt++;
        }
    }
    return cnt;
}

int trt_camera_params(const float pos[3], float yaw, float pitch, float aperture, float focus, int w, int h,
                      void* cam_out) {
    if (!pos || !cam_out || w <= 0 || h <= 0) return fail(TRT_ERR_ARG, "bad arguments");
    CameraController cam(Vec{pos[0], pos[1], pos[2]}, Vec{0, 0, -1});
    cam.set_angles(yaw, pitch);
    cam.set_lens(aperture, focus);
    CameraParams p = cam.get_params(w, h);
    memset(cam_out, 0, sizeof(p));
    memcpy(cam_out, &p, sizeof(p));
    return 0;
}

int trt_scene_create(int config, const char* asset_dir, int grid, void* out, int cap, char* tex_files, int tex_cap) {
    if (config < 0 || config > 5) return fail(TRT_ERR_ARG, "unknown scene config %d", config);
    Scene s = create_config_scene(config, asset_dir, grid);
    if (out) {
        if ((int)s.objects.size() > cap) return fail(TRT_ERR_ARG, "buffer too small: %zu objects", s.objects.size());
        memcpy(out, s.objects.data(), s.objects.size() * sizeof(Object));
    }
    if (tex_files && tex_cap > 0) {
        std::string j;
        for (size_t i = 0; i < s.texture_files.size(); i++) {
            if (i) j += ';';
            j += s.texture_files[i];
        }
        snprintf(tex_files, tex_cap, "%s", j.c_str());
    }
    return (int)s.objects.size();
}

int trt_load_ppm(const char* filename, int* w, int* h, unsigned char** rgb) {
    if (!filename || !w || !h || !rgb) return fail(TRT_ERR_ARG, "null pointer");
    *rgb = load_ppm(filename, w, h);
    if (!*rgb) return fail(TRT_ERR_IO, "cannot read P6 image %s", filename);
    return 0;
}

// Deterministic stand-in for the reference's missing assets/earth.ppm:
// pixel(x,y) = (x*255/(w-1), y*255/(h-1), (x^y)&255).
int trt_write_ppm_earth(const char* filename, int w, int h) {
    if (!filename || w < 2 || h < 2) return fail(TRT_ERR_ARG, "bad arguments");
    FILE* fp = fopen(filename, "wb");
    if (!fp) return fail(TRT_ERR_IO, "cannot write %s", filename);
    fprintf(fp, "P6\n%d %d\n255\n", w, h);
    std::vector<unsigned char> row((size_t)w * 3);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            row[x * 3 + 0] = (unsigned char)(x * 255 / (w - 1));
            row[x * 3 + 1] = (unsigned char)(y * 255 / (h - 1));
            row[x * 3 + 2] = (unsigned char)((x ^ y) & 255);
        }
        fwrite(row.data(), 1, row.size(), fp);
    }
    fclose(fp);
    return 0;
}

void trt_free(void* p) { free(p); }

int trt_xorwow_init_host(uint64_t seed, uint64_t subsequence, uint32_t out[6]) {
    if (!out) return fail(TRT_ERR_ARG, "null pointer");
    xorwow_init_host(seed, subsequence, out, &out[5]);
    return 0;
}

int trt_xorwow_rowcol_host(uint64_t seed, int w, int row, int col, uint32_t out[6]) {
    if (!out || w <= 0 || row < 0 || col < 0 || col >= w) return fail(TRT_ERR_ARG, "bad arguments");
    uint32_t s[5], lo_v[5], cv[5];
    xorwow_seed_state(seed, s, &out[5]);
    // column part through the two-level tables the device kernel uses: M^col = hi[col >> 6] * lo[col & 63]
    std::vector<Gf2Mat> lo, hi;
    xorwow_build_col_levels(w, lo, hi);
    gf2_matvec(lo[col & 63], s, lo_v);
    gf2_matvec(hi[col >> 6], lo_v, cv);
    Gf2Mat mw, mr;
    gf2_pow(xorwow_subsequence_matrix(), (uint64_t)w, mw);
    gf2_pow(mw, (uint64_t)row, mr);
    gf2_matvec(mr, cv, out);
    return 0;
}

int trt_wide_bvh_host(const void* objects, int n_objects, const void* nodes, int n_nodes, void* wide_nodes,
                      int wide_cap, void* tris, int tri_cap, void* leaf_boxes, int info[4]) {
    if (!objects || !nodes || n_objects <= 0 || n_nodes <= 0 || !info) return fail(TRT_ERR_ARG, "bad arguments");
    WideBvh wb;
    build_wide_bvh((const Object*)objects, n_objects, (const LinearBVHNode*)nodes, n_nodes, wb);
    info[0] = (int)wb.nodes.size();
    info[1] = (int)wb.tris.size();
    info[2] = wb.n_top_prims;
    info[3] = wb.depth;
    if (wide_nodes) {
        if (wide_cap < (int)wb.nodes.size()) return fail(TRT_ERR_ARG, "wide node buffer too small");
        memcpy(wide_nodes, wb.nodes.data(), wb.nodes.size() * sizeof(WideNode));
    }
    if (tris) {
        if (tri_cap < (int)wb.tris.size()) return fail(TRT_ERR_ARG, "triangle buffer too small");
        memcpy(tris, wb.tris.data(), wb.tris.size() * sizeof(TriRecord));
    }
    if (leaf_boxes) memcpy(leaf_boxes, wb.leaf_boxes.data(), (size_t)n_objects * sizeof(LeafBox));
    return 0;
}

}  // extern "C"
