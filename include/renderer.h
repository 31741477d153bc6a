// renderer.h -- the renderer boundary.  Drop-in for the reference's
// include/renderer.h (:35-38 init_scene_data, :57 launch_render_kernel); both are
// thin wrappers over the C ABI in trt_capi.h using a process-global context,
// mirroring the reference's file-scope device globals (src/renderer.cu:15-29).
#pragma once
#include "scene.h"
#include "bvh.h"

// Copies objects / BVH nodes / light indices to the GPU, loads the PPM textures,
// and re-lays the scene out for the sm_100a kernels.  Host vectors stay with the caller.
void init_scene_data(const std::vector<Object>& objects,
                     const std::vector<std::string>& texture_files,
                     const std::vector<LinearBVHNode>& nodes,
                     const std::vector<int>& light_indices);

// Adds ONE sample per pixel to the device buffer `accum_buffer` (w*h Vec, running
// sum) using RNG stream (seed 1984+frame_seed, subsequence = pixel index).
// Asynchronous.  tx,ty are accepted for source compatibility and ignored.
void launch_render_kernel(Vec* accum_buffer, int width, int height, int frame_seed,
                          int tx, int ty, CameraParams cam);

// Batched form: frames first_frame_seed .. first_frame_seed+n_frames-1 in one call.
void launch_render_frames(Vec* accum_buffer, int width, int height, int first_frame_seed,
                          int n_frames, CameraParams cam);
