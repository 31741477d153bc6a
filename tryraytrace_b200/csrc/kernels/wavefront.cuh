// wavefront.cuh -- data layout of the streaming wavefront path tracer and the host-side
// launch interface implemented in wavefront.cu.
//
// A fixed pool of P path slots lives in HBM.  Every iteration runs
//     prepare -> regenerate -> [compact] -> extend (closest hit) -> shade -> shadow (any hit)
// over the pool (prepare/regenerate of the next iteration on a side stream, beside the shadow
// kernel; compaction, drain phase only, on the main stream behind it).  Paths that end are accumulated into the caller's buffer and their slots
// go to a free list; `regenerate` refills those slots with the next camera samples of the
// job, so the extend kernel always sees a full pool until the job drains; in the drain phase the
// live slots are compacted into a prefix of the pool (active_cap).  Shadow rays are
// written IN PLACE (one per slot, a valid flag in sh_d.w): almost every diffuse vertex
// spawns one, so compaction would buy nothing, and the persistent traversal kernels pull
// 32-slot chunks of either ray array with TMA bulk copies.  The next-event contribution
// waits in `pend` until the shadow kernel has had its say (it zeroes pend for occluded
// rays); the next shade pass folds it into the path's radiance.
#pragma once
#include "common.cuh"
#include "xorwow.cuh"

namespace trt {

enum { SLOT_DEAD = 0, SLOT_ACTIVE = 1, SLOT_FINISH = 2 };

// Path slot, SoA: 128 bytes per slot over all arrays.
struct PoolView {
    float4* ray_o;   // origin.xyz, length of the slot's shadow ray (it starts at the same point); for a fresh
                     // camera ray (depth 0) .w carries the pixel index instead (thr / rad are not written yet)
    float4* ray_d;   // direction.xyz, flags (int bits): state | depth << 8 | prev_mode << 16
    float2* hit;     // written by extend: t, hit object id | kTriNoDerive (int bits, -1 = miss); verified by shade
    float4* thr;     // throughput.xyz, pixel index (int bits)
    float4* rad;     // radiance.xyz, unused
    float4* pend;    // next-event contribution of the previous vertex (throughput applied), unused
    float4* sh_d;    // shadow ray of this slot: direction.xyz, valid flag (int bits, 1 = trace it)
    uint4* rng_a;    // XORWOW v0..v3
    uint2* rng_b;    // XORWOW v4, d
    uint32_t* dead_mask;  // render pool only: bit b of word w = slot 32*w+b ended in the last shade pass (k_free_scan
                          // turns the words into the free list the next regenerate consumes)
    int capacity;    // multiple of 256 (512 for the render pool: the shade CTA size)
};

// Device-resident control block (one per context).
struct Control {
    unsigned long long next_sample;    // next camera sample of the job to hand out
    unsigned long long total_samples;  // samples in the job (= frames * pixels)
    int n_free;        // free-list entries appended by the last shade pass
    int n_regen;       // slots to regenerate this iteration
    unsigned long long regen_base;  // first sample index for this iteration's regeneration
    int alive;         // slots that hold a live path
    int cursor_extend, cursor_shadow;  // next chunk of the persistent traversal kernels
    // drain-phase compaction (wavefront.cu k_compact_*): slots [0, active_cap) are the only ones any
    // kernel visits; it shrinks when the job has no samples left and the live paths have halved
    int active_cap, compact_go, compact_new_cap, compact_a, compact_b;
    int refill_ticket;  // blocks of the running k_refill that have finished (the last one closes the iteration's bookkeeping)
    // counters (see trt_counters)
    unsigned long long cnt_samples, cnt_closest, cnt_shadow, cnt_replays, cnt_iterations;
    unsigned long long cnt_nodes, cnt_tris;                // all queries (COUNT builds only)
    unsigned long long cnt_nodes_closest, cnt_tris_closest;  // closest-hit queries only
    unsigned long long cnt_tree_closest, cnt_tree_shadow;    // queries that entered the tree (COUNT builds only)
    // lane-utilisation statistics of the persistent traversal kernels (COUNT builds only; printed by
    // trt_get_counters when TRT_TRAV_STATS is set): [0..4] closest, [8..12] shadow:
    // rounds, sum of lanes holding a ray per round, warp-level node steps, warp-level triangle steps, top-phase chunks
    unsigned long long dbg[16];
};

// Per-job constants handed to the kernels by value.
struct JobParams {
    Camera cam;
    RenderConsts rc;
    int first_frame_seed;  // frame seed of job-local frame 0
    int frame_stride;      // frame seed step between job-local frames
    int seed_base;         // 1984
    int n_frames;          // job-local frame count
    const uint4* row_a;        // h window tables of M^(w*row): words 0-3 of the 640 entries ...
    const uint32_t* row_b;     // ... and word 4 (host/xorwow_tables.h xorwow_window_table)
    const XwColVec* col_vecs;  // n_frames * w entries: M^col * v0(frame)
    float* accum;              // caller's buffer, w*h records of 4 floats (x,y,z,pad)
};

// Phase lengths of the speculative while-while traversal (traverse_fast.cuh).
struct Phases {
    int node_iters;  // node steps per phase at most
    int node_min;    // ... and only while at least this many lanes still have node work
    int tri_min;     // extra triangle steps while at least this many lanes have a triangle waiting
    int ranked_top;  // closest hit: rank the root-level primitives per lane and test nearest first (top_closest_ranked)
};

struct LaunchDims {
    int sms;
    int fast_threads;   // threads of the persistent traversal CTAs (one CTA per SM): 512, 768 or 1024
    int smem_nodes;     // wide nodes staged into shared memory per CTA (top of the tree)
    bool wide_loads;    // big scenes: 256-bit node loads instead of staging (traverse_fast.cuh load_node<true>)
    int refill_below;   // idle lanes are refilled when fewer than this many lanes hold a ray
    int shade_block;    // threads per shade CTA (64 .. 512; the pool capacity is a multiple of 512): larger CTAs, fewer
                        // free-list atomics and barriers waiting on them
    int shade_minb;     // 128-thread shade CTAs: resident CTAs per SM the register allocation aims for (8, 10, 12, 14)
    int regen_block;    // threads per regenerate CTA (it shares SMs with the persistent shadow CTAs)
    int compact_quarters;  // drain phase: compact when live paths <= this many quarters of the visited slots (1..3)
    int finish_below;      // drain tail: paths alive at which k_finish_paths runs the rest of the job to completion (0 = never)
    bool shadow_pair;      // combined kernel: any-hit triangle steps test two triangles of a leaf at once
    bool merged_trace;     // shadow rays of the previous shade pass + closest-hit rays in one persistent launch (needs fused_refill)
    bool fused_refill;     // free scan + bookkeeping + regeneration in one kernel on the main stream (k_refill)
    Phases closest_phases, shadow_phases;
};

// shared-memory bytes the persistent kernels need for a given configuration (0 = unsupported)
size_t wf_fast_smem_bytes(int threads, int smem_nodes, bool shadow);
// largest node count that fits beside the stacks and the ray staging buffers
int wf_fast_max_smem_nodes(int threads, size_t smem_limit);

// ---- launchers (wavefront.cu) --------------------------------------------------------
void wf_init_pool(const PoolView& pool, int* free_list, Control* ctl, cudaStream_t s);
void wf_reset_counters(Control* ctl, cudaStream_t s);
void wf_begin_job(Control* ctl, unsigned long long total_samples, int pool_capacity, cudaStream_t s);
void wf_col_table(const uint32_t* col_pows, int n_col_bits, int w, int first_frame_seed, int frame_stride,
                  int seed_base, int n_frames, XwColVec* out, cudaStream_t s);
// Streams of an iteration: everything runs on `main`; with `overlap` the prepare/regenerate part of
// the next iteration runs on `side`, forked from main after the shade kernel (event `fork`) and
// joined before the next extend kernel (event `join`).  The caller records `fork` on main once
// before the first iteration of a job.
struct IterStreams {
    cudaStream_t main, side;
    cudaEvent_t fork, join;
    bool overlap;
    int mark_mask = 0x3f;  // which of the six timing marks of an iteration are recorded
    // what the host knows from its last completion poll (stale values are safe: active_cap only shrinks
    // and next_sample only grows within a job)
    int visit_cap = 0x7fffffff;  // upper bound of Control::active_cap: sizes the shade grid
    bool samples_left = true;    // false: the job has handed out its last sample, nothing to regenerate
    int finish_below = 0;        // > 0: the drain-tail kernel rides along and takes over once this few paths are alive
    bool mostly_live = true;     // at least half of the visited slots hold a path: shade requests a slot's whole state up front
};
// one wavefront iteration on the streams of `st`; returns the number of kernels it launched
// `marks`, when not null, receives six events: [0] prepare+regenerate [1]  and  [2] extend [3] shade [4] shadow [5]
int wf_iteration(const PoolView& pool, int* free_list, Control* ctl, const SceneDev& sc, const TopPrims& top,
                 const JobParams& job, int traversal, bool count, const LaunchDims& dims, const IterStreams& st,
                 cudaEvent_t* marks = nullptr, int* compact_lists = nullptr);
// one-time opt-in to large dynamic shared memory for the persistent kernels
int wf_configure();

// Parity / test entry points.  In FAST mode they run the production persistent kernels over a
// scratch pool (`scratch`, capacity >= n rounded up to 256; `ctl` is the context's control block).
void wf_trace_primary(const SceneDev& sc, const JobParams& job, int frame_seed, int traversal, int* d_id, float* d_t,
                      float* d_ray, uint32_t* d_fetched, uint32_t* d_entered, uint32_t* d_tris, const TopPrims& top,
                      const PoolView& scratch, Control* ctl, const LaunchDims& dims, cudaStream_t s);
void wf_trace_closest(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_id, float* d_t,
                      const TopPrims& top, const PoolView& scratch, Control* ctl, const LaunchDims& dims,
                      cudaStream_t s);
void wf_trace_shadow(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_occ, const TopPrims& top,
                     const PoolView& scratch, Control* ctl, const LaunchDims& dims, cudaStream_t s);
void wf_rng_states(const JobParams& job, int frame_local, int first_pixel, int n, uint32_t* d_states, cudaStream_t s);
void wf_tonemap(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, cudaStream_t s);

}  // namespace trt
