// pipeline.cpp -- display pipeline worker (reference src/pipeline.cpp:13-164): a worker
// thread waits for a dispatched frame, copies the staging buffer to pinned host memory and
// tone-maps it into ARGB8888; dispatch never blocks (frames are dropped while the worker is busy).
#include "pipeline.h"
#include <cuda_runtime.h>

namespace {

void worker_main(Pipeline* p) {
    for (;;) {
        int frame;
        {
            std::unique_lock<std::mutex> lk(p->mtx);
            p->cv_worker.wait(lk, [p] { return p->quit || p->worker_busy; });
            if (p->quit) return;
            frame = p->current_frame;
        }
        cudaMemcpy(p->h_accum, p->d_staging, p->size_bytes, cudaMemcpyDeviceToHost);
        const long long n = (long long)p->width * p->height;
        const float inv = 1.0f / frame;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) {
            const Vec a = p->h_accum[i] * inv;
            const uint32_t r = (uint32_t)toInt(a.x), g = (uint32_t)toInt(a.y), b = (uint32_t)toInt(a.z);
            p->pixel_buffer[i] = (255u << 24) | (r << 16) | (g << 8) | b;
        }
        {
            std::lock_guard<std::mutex> lk(p->mtx);
            p->frame_ready = true;
            p->worker_busy = false;
        }
    }
}

}  // namespace

void pipeline_init(Pipeline* pipe, Vec* h_accum, Vec* d_staging, uint32_t* pixel_buffer, int w, int h) {
    pipe->h_accum = h_accum;
    pipe->d_staging = d_staging;
    pipe->pixel_buffer = pixel_buffer;
    pipe->width = w;
    pipe->height = h;
    pipe->size_bytes = (size_t)w * h * sizeof(Vec);
    pipe->quit = false;
    pipe->worker_busy = false;
    pipe->frame_ready = false;
    pipe->worker_thread = std::thread(worker_main, pipe);
}

bool pipeline_try_dispatch(Pipeline* pipe, int current_gpu_frame) {
    std::lock_guard<std::mutex> lk(pipe->mtx);
    if (pipe->worker_busy) return false;
    pipe->current_frame = current_gpu_frame;
    pipe->worker_busy = true;
    pipe->cv_worker.notify_one();
    return true;
}

bool pipeline_check_frame_ready(Pipeline* pipe) {
    std::lock_guard<std::mutex> lk(pipe->mtx);
    const bool ready = pipe->frame_ready;
    pipe->frame_ready = false;
    return ready;
}

void pipeline_destroy(Pipeline* pipe) {
    {
        std::lock_guard<std::mutex> lk(pipe->mtx);
        pipe->quit = true;
        pipe->worker_busy = true;
    }
    pipe->cv_worker.notify_all();
    if (pipe->worker_thread.joinable()) pipe->worker_thread.join();
}
