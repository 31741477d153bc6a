// wavefront.cuh -- data layout of the streaming wavefront path tracer and the host-side
// launch interface implemented in wavefront.cu.
//
// A fixed pool of P path slots lives in HBM.  Every iteration runs
//     refill -> [compact] -> trace (shadow rays of the previous vertex, then closest hit) -> shade
// over the pool, three launches on one stream.  Paths that end are accumulated into the caller's buffer and
// flagged in a dead mask; `refill` scans that mask and starts the next camera samples of the job in those
// slots, so the traversal always sees a full pool until the job drains; in the drain phase the live slots
// are compacted into a prefix of the pool (active_cap), and the last few ten thousand paths are run to
// completion by one kernel (k_finish_paths).  Shadow rays are written IN PLACE (one per slot: direction in
// sh_d, length in od[2s].w): almost every diffuse vertex spawns one, so compaction would buy nothing, and
// the persistent traversal kernel pulls 32-slot chunks of the ray arrays with TMA bulk copies.  The
// next-event contribution waits in `pend` until the shadow ray has been traced (an occluded ray zeroes it);
// the next shade pass folds it into the path's radiance.
#pragma once
#include "common.cuh"
#include "xorwow.cuh"

namespace trt {

enum { SLOT_DEAD = 0, SLOT_ACTIVE = 1, SLOT_FINISH = 2 };

// Path slot.  Arrays that are always read and written together are paired into 32-byte records, so that a record
// is one DRAM sector: the refill kernel writes scattered slots, and a 16-byte store into a 32-byte sector costs
// a read-modify-write of the sector at eviction.  120 bytes per slot over all arrays.
struct PoolView {
    float4* od;      // 2 float4 per slot.  [2s]: origin.xyz; .w = length of the slot's shadow ray (it starts at the same
                     // point; > 0 = trace it, 0 = none); for a fresh camera ray (depth 0) .w carries the pixel index with
                     // the sign bit set instead (never > 0; thr is not written yet).  [2s+1]: direction.xyz, flags (int
                     // bits): state | depth << 8 | prev_mode << 16
    uint4* rs;       // 2 uint4 per slot.  [2s]: XORWOW v0..v3.  [2s+1]: v4, d, radiance.x, radiance.y (float bits)
    float2* hit;     // written by extend: t, hit object id | kTriNoDerive (int bits, -1 = miss); verified by shade
    float4* thr;     // throughput.xyz, pixel index (int bits)
    float4* pend;    // next-event contribution of the previous vertex (throughput applied).xyz; .w = radiance.z
    float4* sh_d;    // shadow ray of this slot: direction.xyz (valid when od[2s].w > 0), unused
    uint32_t* dead_mask;  // render pool only: bit b of word w = slot 32*w+b ended in the last shade pass (k_refill
                          // turns the words into the slots the next camera samples go to)
    int capacity;    // multiple of 256 (512 for the render pool)
};
constexpr int kFreshPixelBit = (int)0x80000000;

// Device-resident control block (one per context).
struct Control {
    unsigned long long next_sample;    // next camera sample of the job to hand out
    unsigned long long total_samples;  // samples in the job (= frames * pixels)
    int n_free;        // unused
    int n_regen;       // camera samples the last refill started
    unsigned long long regen_base;  // next_sample after the previous refill
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
    unsigned long long cnt_violations;  // LaunchDims::debug_checks: slot-state invariants found broken (must stay 0)
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
    int shade_minb;     // 128-thread shade CTAs: resident CTAs per SM the register allocation aims for (8, 9, 10, 12, 14)
    bool debug_checks;     // k_refill / k_compact_move verify the state of every slot they overwrite (Control::cnt_violations)
    int compact_quarters;  // drain phase: compact when live paths <= this many quarters of the visited slots (1..3)
    int finish_below;      // drain tail: paths alive at which k_finish_paths runs the rest of the job to completion (0 = never)
    bool merged_trace;     // shadow rays of the previous shade pass + closest-hit rays in one persistent launch
    Phases closest_phases, shadow_phases;
};

// shared-memory bytes the persistent kernels need for a given configuration (0 = unsupported)
size_t wf_fast_smem_bytes(int threads, int smem_nodes, bool shadow);
// largest node count that fits beside the stacks and the ray staging buffers
int wf_fast_max_smem_nodes(int threads, size_t smem_limit);

// ---- launchers (wavefront.cu) --------------------------------------------------------
void wf_init_pool(const PoolView& pool, cudaStream_t s);
void wf_reset_counters(Control* ctl, cudaStream_t s);
void wf_begin_job(Control* ctl, unsigned long long total_samples, int pool_capacity, cudaStream_t s);
void wf_col_table(const uint4* col_a, const uint32_t* col_b, int w, int first_frame_seed, int frame_stride,
                  int seed_base, int n_frames, XwColVec* out, cudaStream_t s);
// What the host knows when it issues an iteration (from its last completion poll; stale values are safe:
// active_cap only shrinks and next_sample only grows within a job).
struct IterStreams {
    cudaStream_t main;
    int mark_mask = 0x3f;  // which of the six timing marks of an iteration are recorded
    int visit_cap = 0x7fffffff;  // upper bound of Control::active_cap: sizes the shade and refill grids
    bool samples_left = true;    // false: the job has handed out its last sample, nothing to regenerate
    int finish_below = 0;        // > 0: the drain-tail kernel rides along and takes over once this few paths are alive
    bool mostly_live = true;     // at least half of the visited slots hold a path: shade requests a slot's whole state up front
};
// one wavefront iteration on the streams of `st`; returns the number of kernels it launched
// `marks`, when not null, receives six events: [0] refill [1]  and  [2] traversal [3] shade [4] shadow (REF mode only) [5]
int wf_iteration(const PoolView& pool, Control* ctl, const SceneDev& sc, const TopPrims& top,
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
void wf_compress_nodes(const float4* d_wide, int n, uint4* d_cnodes, int* d_bad, cudaStream_t s);
// d_inst: n_inst records (offset.xyz, scale); writes n_unit * n_inst objects of 7 float4
void wf_instance_objects(const float4* d_unit, int n_unit, const float4* d_inst, int n_inst, float4* d_out, cudaStream_t s);
void wf_tonemap(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, cudaStream_t s);

}  // namespace trt
