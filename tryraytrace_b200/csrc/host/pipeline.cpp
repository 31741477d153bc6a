// pipeline.cpp -- display pipeline worker (reference src/pipeline.cpp:13-164): a worker
// thread waits for a dispatched frame and turns the staging snapshot into the ARGB8888 image the
// window shows; dispatch never blocks (frames are dropped while the worker is busy).
//
// The reference copies the whole float buffer to the host (16 B per pixel) and tone-maps it with
// OpenMP (src/pipeline.cpp:45-71).  Here the tone map runs on the device, on the worker's own
// stream (kernels/wavefront.cu k_tonemap = common.h toInt per channel), and only the 4-byte pixels
// cross PCIe before the frame is reported ready.  The float copy into h_accum, which the reference's
// snapshot key reads (src/main.cpp:161, :224), follows behind it and can be switched off with
// pipeline_set_host_accum(pipe, false) by a caller that does not read h_accum.
#include "pipeline.h"
#include "trt_capi.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <map>

namespace {

// Per-pipeline state that has no place in the caller-allocated struct (its layout is the reference's).
struct Extra {
    uint32_t* d_argb = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
    bool host_accum = true;
    unsigned long long d2h_bytes = 0;
    unsigned long long frames_done = 0;
};
std::mutex g_extra_mtx;
std::map<const Pipeline*, Extra*> g_extra;

Extra* extra_of(const Pipeline* p) {
    std::lock_guard<std::mutex> lk(g_extra_mtx);
    auto it = g_extra.find(p);
    return it == g_extra.end() ? nullptr : it->second;
}

bool cu_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    std::fprintf(stderr, "[Pipeline Error] %s: %s\n", what, cudaGetErrorString(e));
    return false;
}

void worker_main(Pipeline* p, Extra* x) {
    cudaSetDevice(x->device);
    const size_t n = (size_t)p->width * p->height;
    for (;;) {
        int frame;
        bool want_accum;
        {
            std::unique_lock<std::mutex> lk(p->mtx);
            p->cv_worker.wait(lk, [p] { return p->quit || p->worker_busy; });
            if (p->quit) return;
            frame = p->current_frame;
            want_accum = x->host_accum && p->h_accum != nullptr;
        }
        // tone map where the data is, then 4 bytes per pixel to the host
        bool ok = frame > 0 && trt_tonemap_stream(reinterpret_cast<const float*>(p->d_staging), (int)n, frame, x->d_argb, x->stream) == 0;
        ok = ok && cu_ok(cudaMemcpyAsync(p->pixel_buffer, x->d_argb, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, x->stream), "pixel copy");
        ok = ok && cu_ok(cudaStreamSynchronize(x->stream), "tone map");
        {
            std::lock_guard<std::mutex> lk(p->mtx);
            if (ok) {
                p->frame_ready = true;
                x->d2h_bytes += n * sizeof(uint32_t);
                x->frames_done++;
            }
            if (!want_accum) p->worker_busy = false;
        }
        if (!want_accum) continue;
        // the float image for the snapshot key, behind the displayed frame; the worker stays busy, so the
        // staging buffer is not dispatched again before the copy is through
        const bool copied = cu_ok(cudaMemcpyAsync(p->h_accum, p->d_staging, p->size_bytes, cudaMemcpyDeviceToHost, x->stream), "accum copy") &&
                            cu_ok(cudaStreamSynchronize(x->stream), "accum copy");
        std::lock_guard<std::mutex> lk(p->mtx);
        if (copied) x->d2h_bytes += p->size_bytes;
        p->worker_busy = false;
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
    Extra* x = new Extra;
    cudaGetDevice(&x->device);  // the staging buffer lives on the caller's current device
    cu_ok(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking), "stream");
    cu_ok(cudaMalloc(&x->d_argb, (size_t)w * h * sizeof(uint32_t)), "ARGB buffer");
    {
        std::lock_guard<std::mutex> lk(g_extra_mtx);
        g_extra[pipe] = x;
    }
    pipe->worker_thread = std::thread(worker_main, pipe, x);
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
    Extra* x = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_extra_mtx);
        auto it = g_extra.find(pipe);
        if (it != g_extra.end()) {
            x = it->second;
            g_extra.erase(it);
        }
    }
    if (x) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(x->device);
        cudaFree(x->d_argb);
        if (x->stream) cudaStreamDestroy(x->stream);
        cudaSetDevice(prev);
        delete x;
    }
}

void pipeline_set_host_accum(Pipeline* pipe, bool enabled) {
    Extra* x = extra_of(pipe);
    if (!x) return;
    std::lock_guard<std::mutex> lk(pipe->mtx);
    x->host_accum = enabled;
}

unsigned long long pipeline_d2h_bytes(Pipeline* pipe) {
    Extra* x = extra_of(pipe);
    if (!x) return 0;
    std::lock_guard<std::mutex> lk(pipe->mtx);
    return x->d2h_bytes;
}

unsigned long long pipeline_frames_done(Pipeline* pipe) {
    Extra* x = extra_of(pipe);
    if (!x) return 0;
    std::lock_guard<std::mutex> lk(pipe->mtx);
    return x->frames_done;
}
