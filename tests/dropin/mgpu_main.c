/* mgpu_main.c -- plain C host of the single-process multi-GPU layer (include/trt_mgpu.h): builds a
 * config scene through the C ABI, renders one progressive pass on N GPUs (sample-index split + one
 * ncclAllReduce) and the same pass on one GPU, and compares the two accumulation buffers.
 * usage: mgpu_main <asset_dir> <config> <width> <height> <frames> <n_gpus>   -> prints one JSON line */
#include "trt_capi.h"
#include "trt_mgpu.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const char* assets = argv[1];
    const int config = atoi(argv[2]), w = atoi(argv[3]), h = atoi(argv[4]), frames = atoi(argv[5]), n_gpus = atoi(argv[6]);
    char tex[4096];
    int n = trt_scene_create(config, assets, 0, NULL, 0, tex, sizeof tex);
    if (n <= 0) return 3;
    void* objects = malloc((size_t)n * 112);
    void* nodes = malloc((size_t)n * 2 * 48);
    int* lights = (int*)malloc((size_t)n * sizeof(int));
    n = trt_scene_create(config, assets, 0, objects, n, tex, sizeof tex);
    const int nn = trt_bvh_build(objects, n, nodes, 2 * n);
    const int nl = trt_collect_lights(objects, n, lights, n);
    const float pos[3] = {50.f, 45.f, 230.f};
    unsigned char cam[80];
    trt_camera_params(pos, -90.f, -6.f, 0.f, 240.f, w, h, cam);
    const size_t px = (size_t)w * h;
    float* multi = (float*)malloc(px * 16);
    float* single = (float*)malloc(px * 16);

    trt_mgpu* m = NULL;
    if (trt_mgpu_create(n_gpus, NULL, &m) != 0) return 4;
    if (trt_mgpu_upload_scene(m, objects, n, nodes, nn, lights, nl, NULL, 0) != 0) return 5;
    float pass_ms = 0.f;
    if (trt_mgpu_render_to_host(m, multi, w, h, 1, frames, cam, NULL, &pass_ms) != 0) return 6;  /* warm-up */
    if (trt_mgpu_render_to_host(m, multi, w, h, 1, frames, cam, NULL, &pass_ms) != 0) return 6;
    unsigned long long closest = 0, shadow = 0;
    trt_mgpu_rays(m, (uint64_t*)&closest, (uint64_t*)&shadow);
    trt_mgpu_destroy(m);

    trt_ctx* c = NULL;
    if (trt_create(0, &c) != 0) return 7;
    if (trt_upload_scene(c, objects, n, nodes, nn, lights, nl, NULL, 0) != 0) return 8;
    trt_render_to_host(c, single, w, h, 1, frames, 1, cam, NULL);
    const double t0 = now_ms();
    if (trt_render_to_host(c, single, w, h, 1, frames, 1, cam, NULL) != 0) return 9;
    const double single_ms = now_ms() - t0;
    trt_destroy(c);

    double max_rel = 0.0, sum_m = 0.0, sum_s = 0.0;
    for (size_t i = 0; i < px; i++)
        for (int k = 0; k < 3; k++) {
            const double a = multi[i * 4 + k], b = single[i * 4 + k];
            const double rel = fabs(a - b) / fmax(fabs(b), 1.0);
            if (rel > max_rel) max_rel = rel;
            sum_m += a;
            sum_s += b;
        }
    printf("{\"n_gpus\": %d, \"frames\": %d, \"width\": %d, \"height\": %d, \"multi_ms\": %.3f, \"single_ms\": %.3f, "
           "\"multi_mrays_per_s\": %.1f, \"rays\": %llu, \"max_rel_diff\": %.3e, \"sum_multi\": %.6e, \"sum_single\": %.6e}\n",
           n_gpus, frames, w, h, pass_ms, single_ms, (closest + shadow) / pass_ms / 1e3, closest + shadow, max_rel, sum_m, sum_s);
    return 0;
}
