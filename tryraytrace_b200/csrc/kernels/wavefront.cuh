// wavefront.cuh -- data layout of the streaming wavefront path tracer and the host-side
// launch interface implemented in wavefront.cu.
//
// A fixed pool of P path slots lives in HBM (sized so its working set stays L2 resident).
// Every iteration runs
//     prepare -> regenerate -> extend (closest hit) -> shade -> shadow (any hit)
// over the pool.  Paths that end are accumulated into the caller's buffer and their slots
// go to a free list; `regenerate` refills those slots with the next camera samples of the
// job, so the extend kernel always sees a full pool until the job drains.  Shadow rays are
// compacted into their own queue with warp-aggregated atomics (__ballot_sync/__popc).
#pragma once
#include "common.cuh"
#include "xorwow.cuh"

namespace trt {

enum { SLOT_DEAD = 0, SLOT_ACTIVE = 1, SLOT_FINISH = 2 };

// Path slot, SoA: 96 bytes per slot over all arrays.
struct PoolView {
    float4* ray_o;   // origin.xyz, unused
    float4* ray_d;   // direction.xyz, flags (int bits): state | depth << 8 | prev_mode << 16
    float2* hit;     // written by extend: t, hit object id (int bits, -1 = miss)
    float4* thr;     // throughput.xyz, pixel index (int bits)
    float4* rad;     // radiance.xyz, unused
    uint4* rng_a;    // XORWOW v0..v3
    uint2* rng_b;    // XORWOW v4, d
    int capacity;
};

// Shadow-ray queue entry, SoA: 48 bytes.
struct ShadowView {
    float4* o;  // origin.xyz, max_dist
    float4* d;  // direction.xyz, slot (int bits)
    float4* c;  // contribution.rgb (throughput already applied)
};

// Device-resident control block (one per context).
struct Control {
    unsigned long long next_sample;    // next camera sample of the job to hand out
    unsigned long long total_samples;  // samples in the job (= frames * pixels)
    int n_free;        // free-list entries appended by the last shade pass
    int n_regen;       // slots to regenerate this iteration
    unsigned long long regen_base;  // first sample index for this iteration's regeneration
    int n_shadow;      // shadow-queue entries appended by the last shade pass
    int alive;         // slots that hold a live path
    int cursor_extend, cursor_shadow, cursor_replay;
    int n_replay;
    // counters (see trt_counters)
    unsigned long long cnt_samples, cnt_closest, cnt_shadow, cnt_replays, cnt_iterations;
    unsigned long long cnt_nodes, cnt_tris;                // all queries (COUNT builds only)
    unsigned long long cnt_nodes_closest, cnt_tris_closest;  // closest-hit queries only
};

// Per-job constants handed to the kernels by value.
struct JobParams {
    Camera cam;
    RenderConsts rc;
    int first_frame_seed;  // frame seed of job-local frame 0
    int frame_stride;      // frame seed step between job-local frames
    int seed_base;         // 1984
    int n_frames;          // job-local frame count
    const uint32_t* row_mats;  // h matrices, kXwMatWords words each: M^(w*row)
    const XwColVec* col_vecs;  // n_frames * w entries: M^col * v0(frame)
    float* accum;              // caller's buffer, w*h records of 4 floats (x,y,z,pad)
};

struct LaunchDims {
    int sms;
    int fast_blocks_per_sm;  // resident CTAs per SM of the persistent traversal kernels
    int fast_variant;        // 4, 6 or 8: __launch_bounds__ min-blocks variant (register budget)
};

// ---- launchers (wavefront.cu) --------------------------------------------------------
void wf_init_pool(const PoolView& pool, int* free_list, Control* ctl, cudaStream_t s);
void wf_reset_counters(Control* ctl, cudaStream_t s);
void wf_begin_job(Control* ctl, unsigned long long total_samples, int pool_capacity, cudaStream_t s);
void wf_col_table(const uint32_t* col_pows, int n_col_bits, int w, int first_frame_seed, int frame_stride,
                  int seed_base, int n_frames, XwColVec* out, cudaStream_t s);
// one wavefront iteration (five kernels) on stream s
// `marks`, when not null, receives five events recorded at the kernel boundaries of the
// iteration: [prepare+regenerate] m1 [extend] m2 [shade] m3 [shadow] m4  (m0 first).
void wf_iteration(const PoolView& pool, const ShadowView& sq, int* free_list, int* replay_list, Control* ctl,
                  const SceneDev& sc, const JobParams& job, int traversal, bool count, const LaunchDims& dims,
                  cudaStream_t s, cudaEvent_t* marks = nullptr);
int wf_kernels_per_iteration(int traversal);

void wf_trace_primary(const SceneDev& sc, const JobParams& job, int frame_seed, int traversal, int* d_id, float* d_t,
                      float* d_ray, uint32_t* d_fetched, uint32_t* d_entered, uint32_t* d_tris, cudaStream_t s);
void wf_trace_closest(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_id, float* d_t,
                      cudaStream_t s);
void wf_trace_shadow(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_occ, cudaStream_t s);
void wf_rng_states(const JobParams& job, int frame_local, int first_pixel, int n, uint32_t* d_states, cudaStream_t s);
void wf_tonemap(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, cudaStream_t s);

}  // namespace trt
