/* trt_mgpu.h -- single-process multi-GPU rendering for C / C++ hosts (libtrt_b200_mgpu.so).
 *
 * North-star subsystem (3): the progressive pass is partitioned over the GPUs of one box by SAMPLE
 * INDEX -- GPU g of G renders frame seeds first+g, first+g+G, ... with a scene replica of its own,
 * so the union of RNG streams equals the single-GPU run -- and the accumulation buffers are summed
 * with ONE ncclAllReduce per pass over NVLink / NVSwitch.  That collective takes the place of the
 * reference's per-frame device-to-device snapshot (reference src/main.cpp:188): every GPU ends
 * the pass holding the full image.
 *
 * This is the C-ABI twin of tryraytrace_b200/sharding.py (one process per GPU, torch.distributed),
 * for hosts that are not Python.  It lives in its own shared library because it links NCCL, which
 * the one-process-per-GPU harness must not load a second copy of.
 */
#ifndef TRT_MGPU_H
#define TRT_MGPU_H
#include "trt_capi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trt_mgpu trt_mgpu;

/* devices == NULL: GPUs 0 .. n_gpus-1.  Creates one trt_ctx and one NCCL communicator per GPU. */
int trt_mgpu_create(int n_gpus, const int* devices, trt_mgpu** out);
int trt_mgpu_destroy(trt_mgpu* m);
int trt_mgpu_count(const trt_mgpu* m);

/* init_scene_data on every GPU (same arguments as trt_upload_scene). */
int trt_mgpu_upload_scene(trt_mgpu* m, const void* objects, int n_objects, const void* nodes, int n_nodes,
                          const int* lights, int n_lights, const trt_image* textures, int n_textures);

/* One progressive pass of n_frames frames (seeds first .. first+n_frames-1), split over the GPUs by
 * sample index, summed with one all-reduce; the w*h*16-byte result is copied from GPU 0 into h_accum
 * (host memory).  pass_ms, when not NULL, receives the wall time of render + all-reduce. */
int trt_mgpu_render_to_host(trt_mgpu* m, float* h_accum, int width, int height, int first_frame_seed,
                            int n_frames, const void* cam, const trt_opts* opts, float* pass_ms);

/* The same pass ADDED into a caller-owned device buffer on the first GPU of the set (w*h*16 bytes, running sum) --
 * what n_frames calls of launch_render_kernel do to d_accum in the reference's main loop (src/main.cpp:181), spread
 * over the GPUs.  GPU 0 renders its share of the frames straight into d_accum, the others into zeroed buffers of their
 * own, and ONE ncclReduce to GPU 0 (in place on d_accum) sums them: d_accum += all n_frames samples.  This is the
 * entry the drop-in renderer boundary (include/renderer.h launch_render_frames) uses when TRT_GPUS > 1.  d_accum
 * must have been allocated on that first GPU; work already queued on its legacy default stream is waited for. */
int trt_mgpu_render_accumulate(trt_mgpu* m, float* d_accum, int width, int height, int first_frame_seed,
                               int n_frames, const void* cam, const trt_opts* opts, float* pass_ms);

/* closest-hit + shadow queries of the last pass, summed over the GPUs */
int trt_mgpu_rays(trt_mgpu* m, uint64_t* closest, uint64_t* shadow);

#ifdef __cplusplus
}
#endif
#endif /* TRT_MGPU_H */
