// wavefront.cu -- kernels of the streaming wavefront path tracer (see wavefront.cuh).
//
// Replaces the reference megakernel render_kernel_impl (reference src/renderer.cu:317-760):
// regenerate = :319-356 (RNG seeding + primary ray), extend = :371-425, shade = :434-733 and
// :739-759, shadow = :273-314/:692-709.
#include "wavefront.cuh"
#include "raygen.cuh"
#include "shade.cuh"
#include "traverse_ref.cuh"
#include "traverse_wide.cuh"
#include "trt_capi.h"

namespace trt {

namespace {

constexpr int kBlock = 256;

TRT_DEV int pack_flags(int state, int depth, int mode) { return state | (depth << 8) | (mode << 16); }

// Block-level queue append: every thread of the block calls this (converged); threads with
// `want` get a distinct index in [old counter, old counter + n).  One global atomic per block.
TRT_DEV int block_append(bool want, int* counter, int* smem_scratch /* >= 2 + warps ints */) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) smem_scratch[2 + warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            const int c = smem_scratch[2 + w];
            smem_scratch[2 + w] = total;
            total += c;
        }
        smem_scratch[0] = total ? atomicAdd(counter, total) : 0;
    }
    __syncthreads();
    const int idx = smem_scratch[0] + smem_scratch[2 + warp] + rank;
    __syncthreads();  // scratch is reused by the next append
    return idx;
}

// ---- prepare: single thread, advances the queue bookkeeping between iterations --------
__global__ void k_prepare(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n_free = ctl->n_free;
    const unsigned long long remaining = ctl->total_samples - ctl->next_sample;
    const int n_regen = remaining < (unsigned long long)n_free ? (int)remaining : n_free;
    ctl->regen_base = ctl->next_sample;
    ctl->n_regen = n_regen;
    ctl->next_sample += (unsigned long long)n_regen;
    ctl->alive += n_regen - n_free;
    ctl->cnt_samples += (unsigned long long)n_regen;
    ctl->cnt_closest += (unsigned long long)n_regen;
    ctl->cnt_shadow += (unsigned long long)ctl->n_shadow;
    ctl->cnt_iterations += 1;
    ctl->n_free = 0;
    ctl->n_shadow = 0;
    ctl->n_replay = 0;
    ctl->cursor_extend = 0;
    ctl->cursor_shadow = 0;
    ctl->cursor_replay = 0;
}

__global__ void k_begin_job(Control* ctl, unsigned long long total, int capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->next_sample = 0;
    ctl->total_samples = total;
    ctl->n_free = capacity;  // every slot is free (the free list is the identity)
    ctl->alive = capacity;   // prepare subtracts n_free and adds n_regen
    ctl->n_shadow = 0;
    ctl->n_regen = 0;
    ctl->n_replay = 0;
}

__global__ void k_reset_counters(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->cnt_samples = ctl->cnt_closest = ctl->cnt_shadow = ctl->cnt_nodes = ctl->cnt_tris = 0;
    ctl->cnt_replays = ctl->cnt_iterations = 0;
    ctl->cnt_nodes_closest = ctl->cnt_tris_closest = 0;
}

__global__ void k_init_pool(PoolView pool, int* free_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.capacity) return;
    pool.ray_d[i] = make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC)));
    free_list[i] = i;
}

// ---- XORWOW column table: col_vecs[f*w + col] = M^col * v0(frame f) --------------------
__global__ void k_col_table(const uint32_t* __restrict__ col_pows, int n_col_bits, int w, int first_frame_seed,
                            int frame_stride, int seed_base, int n_frames, XwColVec* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_frames * w) return;
    const int f = (int)(i / w), col = (int)(i % w);
    uint32_t v[5], d;
    xw_seed((uint32_t)(seed_base + first_frame_seed + f * frame_stride), v, &d);
    for (int j = 0; j < n_col_bits; j++) {
        if ((col >> j) & 1) {
            uint32_t t[5];
            xw_matvec(col_pows + (size_t)j * kXwMatWords, v, t);
#pragma unroll
            for (int k = 0; k < 5; k++) v[k] = t[k];
        }
    }
    XwColVec e;
#pragma unroll
    for (int k = 0; k < 5; k++) e.v[k] = v[k];
    e.d = d;
    e.pad[0] = e.pad[1] = 0;
    out[i] = e;
}

// RNG state of (job-local frame f, pixel) = row_mats[row] * col_vecs[f][col]
TRT_DEV Xorwow sample_rng(const JobParams& job, int f, int row, int col) {
    const XwColVec* cv = job.col_vecs + (size_t)f * job.rc.width + col;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(cv));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(cv) + 2);
    const uint32_t in[5] = {a.x, a.y, a.z, a.w, b.x};
    uint32_t o[5];
    xw_matvec(job.row_mats + (size_t)row * kXwMatWords, in, o);
    Xorwow s;
    s.v0 = o[0]; s.v1 = o[1]; s.v2 = o[2]; s.v3 = o[3]; s.v4 = o[4];
    s.d = b.y;
    return s;
}

// ---- regenerate: refill freed slots with the next camera samples ----------------------
__global__ void __launch_bounds__(kBlock) k_regen(PoolView pool, const int* __restrict__ free_list,
                                                  const Control* __restrict__ ctl, JobParams job) {
    const int n = ctl->n_regen;
    const unsigned long long base = ctl->regen_base;
    const unsigned pixels = (unsigned)(job.rc.width * job.rc.height);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int slot = free_list[j];
        const unsigned long long s = base + (unsigned long long)j;
        const int f = (int)(s / pixels);
        const int pix = (int)(s % pixels);              // reference pixel index i
        const int row = pix / job.rc.width, col = pix - row * job.rc.width;
        const int y = job.rc.height - 1 - row;          // i = (h-1-y)*w + x  (reference :322)
        Xorwow rng = sample_rng(job, f, row, col);
        const Ray r = primary_ray(job.cam, col, y, job.rc.width, job.rc.height, rng);
        pool.ray_o[slot] = make_float4(r.o.x, r.o.y, r.o.z, 0.f);
        pool.ray_d[slot] = make_float4(r.d.x, r.d.y, r.d.z, i2f(pack_flags(SLOT_ACTIVE, 0, MODE_SPEC)));
        pool.thr[slot] = make_float4(1.f, 1.f, 1.f, i2f(pix));
        pool.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        pool.rng_a[slot] = make_uint4(rng.v0, rng.v1, rng.v2, rng.v3);
        pool.rng_b[slot] = make_uint2(rng.v4, rng.d);
    }
}

// ---- extend: closest hit for every active slot -----------------------------------------
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kBlock) k_extend(PoolView pool, SceneDev sc, Control* ctl, int* replay_list) {
    unsigned long long nodes = 0, tris = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < pool.capacity; slot += gridDim.x * blockDim.x) {
        const float4 d4 = pool.ray_d[slot];
        if ((f2i(d4.w) & 0xff) != SLOT_ACTIVE) continue;
        const float4 o4 = pool.ray_o[slot];
        Ray r;
        r.o = f3(o4.x, o4.y, o4.z);
        r.d = f3(d4.x, d4.y, d4.z);
        float t;
        int id;
        if (MODE == TRT_TRAVERSE_REF) {
            VisitCounts vc = {0, 0, 0};
            id = ref_closest<COUNT>(sc, r, &t, &vc);
            if (COUNT) { nodes += vc.fetched; tris += vc.tris; }
        } else {
            WideCounts wc = {0, 0};
            bool ambiguous;
            id = wide_closest<COUNT>(sc, r, &t, &ambiguous, &wc);
            if (COUNT) { nodes += wc.nodes; tris += wc.tris; }
            if (ambiguous) {  // rare: re-run in reference order
                VisitCounts vc = {0, 0, 0};
                id = ref_closest<COUNT>(sc, r, &t, &vc);
                if (COUNT) { nodes += vc.fetched; tris += vc.tris; }
                atomicAdd(&ctl->cnt_replays, 1ull);
            }
        }
        pool.hit[slot] = make_float2(t, i2f(id));
    }
    if (COUNT) {
        atomicAdd(&ctl->cnt_nodes, nodes);
        atomicAdd(&ctl->cnt_tris, tris);
        atomicAdd(&ctl->cnt_nodes_closest, nodes);
        atomicAdd(&ctl->cnt_tris_closest, tris);
    }
    (void)replay_list;
}

// ---- shade: one thread per slot --------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_shade(PoolView pool, ShadowView sq, int* __restrict__ free_list,
                                                  Control* ctl, SceneDev sc, JobParams job) {
    __shared__ int scratch[2 + kBlock / 32];
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;  // capacity is a multiple of kBlock
    const float4 d4 = pool.ray_d[slot];
    const int flags = f2i(d4.w);
    const int state = flags & 0xff;
    bool terminated = false, cont = false;
    PathVertexIO io;
    io.shadow = false;
    int pix = 0;
    if (state != SLOT_DEAD) {
        const float4 thr4 = pool.thr[slot];
        const float4 rad4 = pool.rad[slot];
        pix = f2i(thr4.w);
        io.thr = f3(thr4.x, thr4.y, thr4.z);
        io.rad = f3(rad4.x, rad4.y, rad4.z);
        if (state == SLOT_ACTIVE) {
            const float2 hit = pool.hit[slot];
            const int id = f2i(hit.y);
            if (id < 0) {
                terminated = true;  // miss: black environment (reference :427)
            } else {
                const float4 o4 = pool.ray_o[slot];
                const uint4 ra = pool.rng_a[slot];
                const uint2 rb = pool.rng_b[slot];
                io.ray.o = f3(o4.x, o4.y, o4.z);
                io.ray.d = f3(d4.x, d4.y, d4.z);
                io.depth = (flags >> 8) & 0xff;
                io.prev_mode = (flags >> 16) & 0xff;
                io.rng.v0 = ra.x; io.rng.v1 = ra.y; io.rng.v2 = ra.z; io.rng.v3 = ra.w;
                io.rng.v4 = rb.x; io.rng.d = rb.y;
                if (!shade_vertex(sc, job.rc, io, id, hit.x)) {
                    terminated = true;
                } else {
                    const int depth = io.depth + 1;
                    // the reference loop ends after max_depth vertices; a path that still has a
                    // shadow ray in flight is finalised one iteration later (SLOT_FINISH)
                    const int ns = depth >= job.rc.max_depth ? SLOT_FINISH : SLOT_ACTIVE;
                    cont = ns == SLOT_ACTIVE;
                    pool.ray_o[slot] = make_float4(io.ray.o.x, io.ray.o.y, io.ray.o.z, 0.f);
                    pool.ray_d[slot] = make_float4(io.ray.d.x, io.ray.d.y, io.ray.d.z,
                                                   i2f(pack_flags(ns, depth, io.prev_mode)));
                    pool.thr[slot] = make_float4(io.thr.x, io.thr.y, io.thr.z, thr4.w);
                    pool.rad[slot] = make_float4(io.rad.x, io.rad.y, io.rad.z, 0.f);
                    pool.rng_a[slot] = make_uint4(io.rng.v0, io.rng.v1, io.rng.v2, io.rng.v3);
                    pool.rng_b[slot] = make_uint2(io.rng.v4, io.rng.d);
                }
            }
        } else {
            terminated = true;  // SLOT_FINISH
        }
        if (terminated) {
            F3 rad = io.rad;
            if (filter_sample(rad)) {  // reference :739-759
                float* a = job.accum + (size_t)pix * 4;
                atomicAdd(a + 0, rad.x);
                atomicAdd(a + 1, rad.y);
                atomicAdd(a + 2, rad.z);
            }
            pool.ray_d[slot] = make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC)));
        }
    }
    const int fi = block_append(terminated, &ctl->n_free, scratch);
    if (terminated) free_list[fi] = slot;
    const int si = block_append(io.shadow, &ctl->n_shadow, scratch);
    if (io.shadow) {
        sq.o[si] = make_float4(io.shadow_ray.o.x, io.shadow_ray.o.y, io.shadow_ray.o.z, io.shadow_max_dist);
        sq.d[si] = make_float4(io.shadow_ray.d.x, io.shadow_ray.d.y, io.shadow_ray.d.z, i2f(slot));
        sq.c[si] = make_float4(io.shadow_contrib.x, io.shadow_contrib.y, io.shadow_contrib.z, 0.f);
    }
    // closest-hit rays issued for the next iteration (ray statistics are always maintained)
    const int n_cont = __syncthreads_count(cont);
    if (threadIdx.x == 0 && n_cont) atomicAdd(&ctl->cnt_closest, (unsigned long long)n_cont);
}

// ---- shadow: any hit for every queued shadow ray ----------------------------------------
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kBlock) k_shadow(PoolView pool, ShadowView sq, SceneDev sc, Control* ctl) {
    const int n = ctl->n_shadow;
    unsigned long long nodes = 0, tris = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const float4 o4 = sq.o[j], d4 = sq.d[j];
        Ray r;
        r.o = f3(o4.x, o4.y, o4.z);
        r.d = f3(d4.x, d4.y, d4.z);
        bool occluded;
        if (MODE == TRT_TRAVERSE_REF) {
            VisitCounts vc = {0, 0, 0};
            occluded = ref_shadow<COUNT>(sc, r, o4.w, &vc);
            if (COUNT) { nodes += vc.fetched; tris += vc.tris; }
        } else {
            WideCounts wc = {0, 0};
            occluded = wide_shadow<COUNT>(sc, r, o4.w, &wc);
            if (COUNT) { nodes += wc.nodes; tris += wc.tris; }
        }
        if (!occluded) {
            const int slot = f2i(d4.w);
            const float4 c = sq.c[j];
            float4 rad = pool.rad[slot];  // one shadow ray per slot per iteration: no race
            rad.x += c.x; rad.y += c.y; rad.z += c.z;
            pool.rad[slot] = rad;
        }
    }
    if (COUNT) {
        atomicAdd(&ctl->cnt_nodes, nodes);
        atomicAdd(&ctl->cnt_tris, tris);
    }
}


// ---- persistent fast-path kernels: while-while traversal with dynamic ray fetch -------------
// One warp holds 32 rays.  Rounds (traverse_wide.cuh) keep the lanes converged; a lane whose
// ray has finished writes its result and goes idle; when a quarter of the warp is idle the
// idle lanes grab the next slots from a global cursor with one warp-aggregated atomic
// (__ballot_sync + __popc) and continue -- so a long ray never holds 31 finished lanes hostage.
constexpr int kFastBlock = 128;
constexpr int kRefillBelow = 25;  // refill when fewer than this many lanes hold a ray

template <bool COUNT, int MINB>
__global__ void __launch_bounds__(kFastBlock, MINB) k_extend_fast(PoolView pool, SceneDev sc, Control* ctl) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    ClosestState st;
    ClosestStack stack;
    st.nsp = 0;
    st.tsp = 0;
    st.cur = kWideEmptyRef;
    WideCounts wc = {0, 0};
    bool has = false;
    bool drained = false;  // warp-uniform: the cursor ran past the pool
    int slot = -1;
    unsigned replays = 0;
    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, has);
        if (!drained && __popc(act) < kRefillBelow) {
            const unsigned idle = ~act;
            const int n = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctl->cursor_extend, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!has) {
                const int my = base + __popc(idle & lt_mask);
                if (my < pool.capacity) {
                    const float4 d4 = pool.ray_d[my];
                    if ((f2i(d4.w) & 0xff) == SLOT_ACTIVE) {
                        const float4 o4 = pool.ray_o[my];
                        Ray r;
                        r.o = f3(o4.x, o4.y, o4.z);
                        r.d = f3(d4.x, d4.y, d4.z);
                        closest_begin(st, r);
                        has = true;
                        slot = my;
                    }
                }
            }
            drained = base + n >= pool.capacity;
            act = __ballot_sync(0xffffffffu, has);
        }
        if (act == 0) {
            if (drained) break;
            continue;
        }
        if (has) {
            closest_round<COUNT>(sc, st, stack, &wc);
            if (closest_done(st)) {
                float t = st.d_min;
                int id = st.id;
                if (st.amb) {  // rare: order-dependent reach, re-run in reference order
                    Ray r;
                    r.o = st.o;
                    r.d = st.d;
                    VisitCounts vc = {0, 0, 0};
                    id = ref_closest<false>(sc, r, &t, &vc);
                    replays++;
                }
                pool.hit[slot] = make_float2(t, i2f(id));
                has = false;
            }
        }
    }
    if (COUNT) {
        atomicAdd(&ctl->cnt_nodes, (unsigned long long)wc.nodes);
        atomicAdd(&ctl->cnt_tris, (unsigned long long)wc.tris);
        atomicAdd(&ctl->cnt_nodes_closest, (unsigned long long)wc.nodes);
        atomicAdd(&ctl->cnt_tris_closest, (unsigned long long)wc.tris);
    }
    if (replays) atomicAdd(&ctl->cnt_replays, (unsigned long long)replays);
}

template <bool COUNT, int MINB>
__global__ void __launch_bounds__(kFastBlock, MINB) k_shadow_fast(PoolView pool, ShadowView sq, SceneDev sc, Control* ctl) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_rays = ctl->n_shadow;
    ShadowState st;
    ShadowStack stack;
    st.nsp = 0;
    st.tsp = 0;
    st.occluded = false;
    WideCounts wc = {0, 0};
    bool has = false;
    bool drained = false;
    int entry = -1;
    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, has);
        if (!drained && __popc(act) < kRefillBelow) {
            const unsigned idle = ~act;
            const int n = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctl->cursor_shadow, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!has) {
                const int my = base + __popc(idle & lt_mask);
                if (my < n_rays) {
                    const float4 o4 = sq.o[my], d4 = sq.d[my];
                    Ray r;
                    r.o = f3(o4.x, o4.y, o4.z);
                    r.d = f3(d4.x, d4.y, d4.z);
                    shadow_begin(st, stack, r, o4.w);
                    has = true;
                    entry = my;
                }
            }
            drained = base + n >= n_rays;
            act = __ballot_sync(0xffffffffu, has);
        }
        if (act == 0) {
            if (drained) break;
            continue;
        }
        if (has) {
            shadow_round<COUNT>(sc, st, stack, &wc);
            if (shadow_done(st)) {
                if (!st.occluded) {
                    const int slot = f2i(sq.d[entry].w);
                    const float4 c = sq.c[entry];
                    float4 rad = pool.rad[slot];  // one shadow ray per slot per iteration: no race
                    rad.x += c.x; rad.y += c.y; rad.z += c.z;
                    pool.rad[slot] = rad;
                }
                has = false;
            }
        }
    }
    if (COUNT) {
        atomicAdd(&ctl->cnt_nodes, (unsigned long long)wc.nodes);
        atomicAdd(&ctl->cnt_tris, (unsigned long long)wc.tris);
    }
}

// ---- parity / test entry points -------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace_primary(SceneDev sc, JobParams job, int* out_id, float* out_t,
                                                          float* out_ray, uint32_t* out_fetched,
                                                          uint32_t* out_entered, uint32_t* out_tris) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= job.rc.width * job.rc.height) return;
    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
    const int y = job.rc.height - 1 - row;
    Xorwow rng = sample_rng(job, 0, row, col);
    const Ray r = primary_ray(job.cam, col, y, job.rc.width, job.rc.height, rng);
    float t;
    int id;
    VisitCounts vc = {0, 0, 0};
    if (MODE == TRT_TRAVERSE_REF) {
        id = ref_closest<true>(sc, r, &t, &vc);
    } else {
        WideCounts wc = {0, 0};
        bool ambiguous;
        id = wide_closest<true>(sc, r, &t, &ambiguous, &wc);
        vc.fetched = wc.nodes;
        vc.tris = wc.tris;
        vc.entered = ambiguous ? 1u : 0u;  // FAST mode reports "replayed" here
        if (ambiguous) {
            VisitCounts v2 = {0, 0, 0};
            id = ref_closest<false>(sc, r, &t, &v2);
        }
    }
    if (out_id) out_id[pix] = id;
    if (out_t) out_t[pix] = t;
    if (out_ray) {
        float* p = out_ray + (size_t)pix * 6;
        p[0] = r.o.x; p[1] = r.o.y; p[2] = r.o.z; p[3] = r.d.x; p[4] = r.d.y; p[5] = r.d.z;
    }
    if (out_fetched) out_fetched[pix] = vc.fetched;
    if (out_entered) out_entered[pix] = vc.entered;
    if (out_tris) out_tris[pix] = vc.tris;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace_closest(SceneDev sc, const float* __restrict__ rays, int n,
                                                          int* out_id, float* out_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 8;
    Ray r;
    r.o = f3(p[0], p[1], p[2]);
    r.d = f3(p[3], p[4], p[5]);
    float t;
    int id;
    if (MODE == TRT_TRAVERSE_REF) {
        VisitCounts vc = {0, 0, 0};
        id = ref_closest<false>(sc, r, &t, &vc);
    } else {
        WideCounts wc = {0, 0};
        bool ambiguous;
        id = wide_closest<false>(sc, r, &t, &ambiguous, &wc);
        if (ambiguous) {
            VisitCounts vc = {0, 0, 0};
            id = ref_closest<false>(sc, r, &t, &vc);
        }
    }
    out_id[i] = id;
    if (out_t) out_t[i] = t;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace_shadow(SceneDev sc, const float* __restrict__ rays, int n,
                                                         int* out_occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 8;
    Ray r;
    r.o = f3(p[0], p[1], p[2]);
    r.d = f3(p[3], p[4], p[5]);
    bool occ;
    if (MODE == TRT_TRAVERSE_REF) {
        VisitCounts vc = {0, 0, 0};
        occ = ref_shadow<false>(sc, r, p[6], &vc);
    } else {
        WideCounts wc = {0, 0};
        occ = wide_shadow<false>(sc, r, p[6], &wc);
    }
    out_occ[i] = occ ? 1 : 0;
}

__global__ void k_rng_states(JobParams job, int f, int first_pixel, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pix = first_pixel + i;
    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
    const Xorwow s = sample_rng(job, f, row, col);
    uint32_t* o = out + (size_t)i * 6;
    o[0] = s.v0; o[1] = s.v1; o[2] = s.v2; o[3] = s.v3; o[4] = s.v4; o[5] = s.d;
}

// accum/frames -> gamma 2.2 -> ARGB8888 (reference src/pipeline.cpp:59-71, common.h:114-128)
__global__ void k_tonemap(const float4* __restrict__ accum, int n, float inv_frames, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = accum[i];
    const float c[3] = {a.x * inv_frames, a.y * inv_frames, a.z * inv_frames};
    int q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float v = c[k] < 0.f ? 0.f : (c[k] > 1.f ? 1.f : c[k]);
        q[k] = (int)(pow((double)v, 1 / 2.2) * 255 + .5);
    }
    out[i] = (255u << 24) | ((uint32_t)q[0] << 16) | ((uint32_t)q[1] << 8) | (uint32_t)q[2];
}

int grid_for(int n) { return (n + kBlock - 1) / kBlock; }

}  // namespace

// ---- launchers -------------------------------------------------------------------------------
void wf_init_pool(const PoolView& pool, int* free_list, Control* ctl, cudaStream_t s) {
    (void)ctl;
    k_init_pool<<<grid_for(pool.capacity), kBlock, 0, s>>>(pool, free_list);
}

void wf_reset_counters(Control* ctl, cudaStream_t s) { k_reset_counters<<<1, 32, 0, s>>>(ctl); }

void wf_begin_job(Control* ctl, unsigned long long total_samples, int pool_capacity, cudaStream_t s) {
    k_begin_job<<<1, 32, 0, s>>>(ctl, total_samples, pool_capacity);
}

void wf_col_table(const uint32_t* col_pows, int n_col_bits, int w, int first_frame_seed, int frame_stride,
                  int seed_base, int n_frames, XwColVec* out, cudaStream_t s) {
    const long long n = (long long)n_frames * w;
    k_col_table<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(col_pows, n_col_bits, w, first_frame_seed, frame_stride,
                                                            seed_base, n_frames, out);
}

int wf_kernels_per_iteration(int) { return 5; }

template <int MODE, bool COUNT>
static void iteration_impl(const PoolView& pool, const ShadowView& sq, int* free_list, int* replay_list, Control* ctl,
                           const SceneDev& sc, const JobParams& job, const LaunchDims& dims, cudaStream_t s,
                           cudaEvent_t* marks) {
    const int full = pool.capacity / kBlock;
    const int persistent = dims.sms * 8;
    auto mark = [&](int i) { if (marks) cudaEventRecord(marks[i], s); };
    mark(0);
    k_prepare<<<1, 32, 0, s>>>(ctl);
    k_regen<<<persistent < full ? persistent : full, kBlock, 0, s>>>(pool, free_list, ctl, job);
    mark(1);
    if (MODE == TRT_TRAVERSE_FAST) {
        // persistent grids: enough CTAs to fill every SM, each warp pulls rays until the queue is dry
        const int fast_grid = dims.sms * dims.fast_blocks_per_sm;
        switch (dims.fast_variant) {  // register budget of the persistent kernels (tuning knob)
        case 4:
            k_extend_fast<COUNT, 4><<<fast_grid, kFastBlock, 0, s>>>(pool, sc, ctl);
            mark(2);
            k_shade<COUNT><<<full, kBlock, 0, s>>>(pool, sq, free_list, ctl, sc, job);
            mark(3);
            k_shadow_fast<COUNT, 4><<<fast_grid, kFastBlock, 0, s>>>(pool, sq, sc, ctl);
            break;
        case 6:
            k_extend_fast<COUNT, 6><<<fast_grid, kFastBlock, 0, s>>>(pool, sc, ctl);
            mark(2);
            k_shade<COUNT><<<full, kBlock, 0, s>>>(pool, sq, free_list, ctl, sc, job);
            mark(3);
            k_shadow_fast<COUNT, 6><<<fast_grid, kFastBlock, 0, s>>>(pool, sq, sc, ctl);
            break;
        default:
            k_extend_fast<COUNT, 8><<<fast_grid, kFastBlock, 0, s>>>(pool, sc, ctl);
            mark(2);
            k_shade<COUNT><<<full, kBlock, 0, s>>>(pool, sq, free_list, ctl, sc, job);
            mark(3);
            k_shadow_fast<COUNT, 8><<<fast_grid, kFastBlock, 0, s>>>(pool, sq, sc, ctl);
            break;
        }
    } else {
        k_extend<MODE, COUNT><<<full, kBlock, 0, s>>>(pool, sc, ctl, replay_list);
        mark(2);
        k_shade<COUNT><<<full, kBlock, 0, s>>>(pool, sq, free_list, ctl, sc, job);
        mark(3);
        k_shadow<MODE, COUNT><<<full, kBlock, 0, s>>>(pool, sq, sc, ctl);
    }
    mark(4);
}

void wf_iteration(const PoolView& pool, const ShadowView& sq, int* free_list, int* replay_list, Control* ctl,
                  const SceneDev& sc, const JobParams& job, int traversal, bool count, const LaunchDims& dims,
                  cudaStream_t s, cudaEvent_t* marks) {
    if (traversal == TRT_TRAVERSE_REF) {
        if (count) iteration_impl<TRT_TRAVERSE_REF, true>(pool, sq, free_list, replay_list, ctl, sc, job, dims, s, marks);
        else iteration_impl<TRT_TRAVERSE_REF, false>(pool, sq, free_list, replay_list, ctl, sc, job, dims, s, marks);
    } else {
        if (count) iteration_impl<TRT_TRAVERSE_FAST, true>(pool, sq, free_list, replay_list, ctl, sc, job, dims, s, marks);
        else iteration_impl<TRT_TRAVERSE_FAST, false>(pool, sq, free_list, replay_list, ctl, sc, job, dims, s, marks);
    }
}

void wf_trace_primary(const SceneDev& sc, const JobParams& job, int, int traversal, int* d_id, float* d_t,
                      float* d_ray, uint32_t* d_fetched, uint32_t* d_entered, uint32_t* d_tris, cudaStream_t s) {
    const int n = job.rc.width * job.rc.height;
    if (traversal == TRT_TRAVERSE_REF)
        k_trace_primary<TRT_TRAVERSE_REF><<<grid_for(n), kBlock, 0, s>>>(sc, job, d_id, d_t, d_ray, d_fetched,
                                                                          d_entered, d_tris);
    else
        k_trace_primary<TRT_TRAVERSE_FAST><<<grid_for(n), kBlock, 0, s>>>(sc, job, d_id, d_t, d_ray, d_fetched,
                                                                           d_entered, d_tris);
}

void wf_trace_closest(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_id, float* d_t,
                      cudaStream_t s) {
    if (traversal == TRT_TRAVERSE_REF)
        k_trace_closest<TRT_TRAVERSE_REF><<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_id, d_t);
    else
        k_trace_closest<TRT_TRAVERSE_FAST><<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_id, d_t);
}

void wf_trace_shadow(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_occ, cudaStream_t s) {
    if (traversal == TRT_TRAVERSE_REF)
        k_trace_shadow<TRT_TRAVERSE_REF><<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_occ);
    else
        k_trace_shadow<TRT_TRAVERSE_FAST><<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_occ);
}

void wf_rng_states(const JobParams& job, int frame_local, int first_pixel, int n, uint32_t* d_states,
                   cudaStream_t s) {
    k_rng_states<<<grid_for(n), kBlock, 0, s>>>(job, frame_local, first_pixel, n, d_states);
}

void wf_tonemap(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, cudaStream_t s) {
    k_tonemap<<<grid_for(n_pixels), kBlock, 0, s>>>(reinterpret_cast<const float4*>(d_accum), n_pixels,
                                                    1.0f / frames, d_argb);
}

}  // namespace trt
