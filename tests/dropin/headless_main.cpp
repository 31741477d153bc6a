// headless_main.cpp -- the call sequence of the reference's main loop (reference src/main.cpp:79-116,
// :170-224) without the SDL window, written ONLY against the drop-in headers of include/ and linked
// against libtrt_b200.so: what a TryRaytrace checkout does after the relink of INTEGRATION.md.
//   scene factory -> BVH::build -> light list -> init_scene_data -> cudaMalloc/cudaMemset accum ->
//   per frame: launch_render_kernel, device-to-device snapshot, cudaDeviceSynchronize,
//   pipeline_try_dispatch / pipeline_check_frame_ready -> save raw accumulation + ARGB image.
// All CUDA calls here are on the legacy default stream, exactly like the reference's.
// usage: headless_main <asset_dir> <config> <width> <height> <frames> <out_prefix> [noaccum]
// noaccum: the caller does not read h_accum (pipeline_set_host_accum(false)): 4 bytes per pixel per displayed frame
// batch:   all frames in ONE launch_render_frames call (the batched entry; with TRT_GPUS=N it is split over N GPUs)
#include "bvh.h"
#include "camera.h"
#include "pipeline.h"
#include "renderer.h"
#include "scene.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const char* asset_dir = argv[1];
    const int config = std::atoi(argv[2]), width = std::atoi(argv[3]), height = std::atoi(argv[4]);
    const int frames = std::atoi(argv[5]);
    const std::string prefix = argv[6];

    Scene scene = config == 0 ? create_cornell_box() : create_config_scene(config, asset_dir);
    BVH bvh;
    bvh.build(scene.objects);  // reorders the objects (main.cpp:85)
    std::vector<int> light_indices;  // main.cpp:88-96
    for (size_t i = 0; i < scene.objects.size(); i++) {
        const Vec& e = scene.objects[i].emission;
        if (e.x > 0.1f || e.y > 0.1f || e.z > 0.1f) light_indices.push_back((int)i);
    }
    init_scene_data(scene.objects, scene.texture_files, bvh.get_nodes(), light_indices);  // main.cpp:101

    CameraController cam(make_vec(50.f, 50.f, 295.6f), make_vec(50.f, 50.f, 0.f));  // main.cpp:105
    const size_t n = (size_t)width * height;
    Vec *d_accum = nullptr, *d_staging = nullptr, *h_accum = nullptr;  // main.cpp:110-128
    if (cudaMalloc(&d_accum, n * sizeof(Vec)) != cudaSuccess || cudaMalloc(&d_staging, n * sizeof(Vec)) != cudaSuccess ||
        cudaMallocHost(&h_accum, n * sizeof(Vec)) != cudaSuccess)
        return 3;
    cudaMemset(d_accum, 0, n * sizeof(Vec));
    std::vector<uint32_t> pixels(n);
    Pipeline pipe;
    pipeline_init(&pipe, h_accum, d_staging, pixels.data(), width, height);  // main.cpp:139
    const bool no_accum = argc > 7 && std::string(argv[7]) == "noaccum";
    if (no_accum) pipeline_set_host_accum(&pipe, false);

    const bool batch = argc > 7 && std::string(argv[7]) == "batch";
    int shown = 0;
    if (batch) {
        launch_render_frames(d_accum, width, height, 1, frames, cam.get_params(width, height));
        cudaMemcpy(d_staging, d_accum, n * sizeof(Vec), cudaMemcpyDeviceToDevice);
        cudaDeviceSynchronize();
    }
    for (int gpu_frame = 1; gpu_frame <= frames && !batch; gpu_frame++) {  // main.cpp:152-223
        CameraParams cp = cam.get_params(width, height);
        launch_render_kernel(d_accum, width, height, gpu_frame, 16, 16, cp);          // :181
        cudaMemcpy(d_staging, d_accum, n * sizeof(Vec), cudaMemcpyDeviceToDevice);   // :188
        cudaDeviceSynchronize();                                                      // :192
        pipeline_try_dispatch(&pipe, gpu_frame);                                      // :198
        if (pipeline_check_frame_ready(&pipe)) shown++;                               // :203
    }
    // The worker may still be busy with an earlier snapshot.  A dispatch is accepted only when it is idle, so
    // once a SECOND dispatch of the final snapshot has been accepted the first one is through (pixels and float
    // copy), and once a third has been accepted the ARGB image is certainly the final frame's (all three
    // process the same staging contents).  pipeline_destroy joins the last one.
    for (int round = 0, spin = 0; round < 3; round++)
        while (!pipeline_try_dispatch(&pipe, frames) && spin++ < 20000) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    if (pipeline_check_frame_ready(&pipe)) shown++;
    std::printf("pipeline: frames_done %llu d2h_bytes %llu shown %d\n", pipeline_frames_done(&pipe), pipeline_d2h_bytes(&pipe), shown);
    std::vector<Vec> host(n);
    cudaMemcpy(host.data(), d_accum, n * sizeof(Vec), cudaMemcpyDeviceToHost);
    pipeline_destroy(&pipe);
    FILE* f = std::fopen((prefix + ".accum").c_str(), "wb");
    if (!f) return 4;
    std::fwrite(host.data(), sizeof(Vec), n, f);
    std::fclose(f);
    if (!no_accum) {  // what the snapshot key would save (main.cpp:161, :224)
        f = std::fopen((prefix + ".haccum").c_str(), "wb");
        if (!f) return 4;
        std::fwrite(h_accum, sizeof(Vec), n, f);
        std::fclose(f);
    }
    f = std::fopen((prefix + ".argb").c_str(), "wb");
    if (!f) return 4;
    std::fwrite(pixels.data(), 4, n, f);
    std::fclose(f);
    std::printf("headless ok: %d frames, %zu objects, %zu nodes, %zu lights\n", frames, scene.objects.size(),
                bvh.get_nodes().size(), light_indices.size());
    cudaFree(d_accum);
    cudaFree(d_staging);
    cudaFreeHost(h_accum);
    return 0;
}
