// trt_mgpu.cpp -- see include/trt_mgpu.h.  One host thread per GPU drives that GPU's trt_ctx (the
// render call polls its job to completion, so the GPUs must be driven concurrently); the pass ends
// with one grouped ncclAllReduce of the per-GPU accumulation buffers.
#include "trt_mgpu.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

struct trt_mgpu {
    std::vector<int> devices;
    std::vector<trt_ctx*> ctx;
    std::vector<ncclComm_t> comms;
    std::vector<float*> accum;  // per GPU, w*h*4 floats
    std::vector<cudaStream_t> streams;
    size_t accum_bytes = 0;
};

namespace {
thread_local std::string g_merr;
int mfail(const char* what, const char* detail) {
    g_merr = std::string(what) + ": " + (detail ? detail : "");
    std::fprintf(stderr, "[trt_mgpu] %s\n", g_merr.c_str());
    return TRT_ERR_CUDA;
}
#define MCU(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return mfail(#call, cudaGetErrorString(e_));    \
    } while (0)
// every entry point leaves the calling thread's current device as it found it (the reference application never
// changes it; the drop-in boundary allocates its buffers on it after init_scene_data has returned)
struct DeviceGuard {
    int prev = 0;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};
#define MNCCL(call)                                                            \
    do {                                                                       \
        ncclResult_t r_ = (call);                                              \
        if (r_ != ncclSuccess) return mfail(#call, ncclGetErrorString(r_));    \
    } while (0)
}  // namespace

extern "C" {

int trt_mgpu_create(int n_gpus, const int* devices, trt_mgpu** out) {
    if (!out || n_gpus < 1) return TRT_ERR_ARG;
    DeviceGuard guard;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_gpus) return mfail("trt_mgpu_create", "not enough CUDA devices");
    trt_mgpu* m = new trt_mgpu();
    for (int g = 0; g < n_gpus; g++) m->devices.push_back(devices ? devices[g] : g);
    m->ctx.assign(n_gpus, nullptr);
    m->accum.assign(n_gpus, nullptr);
    m->streams.assign(n_gpus, nullptr);
    m->comms.assign(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; g++) {
        if (int rc = trt_create(m->devices[g], &m->ctx[g])) {
            mfail("trt_create", trt_last_error());
            trt_mgpu_destroy(m);  // the contexts created so far
            return rc;
        }
        m->streams[g] = (cudaStream_t)trt_stream(m->ctx[g]);
    }
    if (ncclResult_t r = ncclCommInitAll(m->comms.data(), n_gpus, m->devices.data()); r != ncclSuccess) {
        mfail("ncclCommInitAll", ncclGetErrorString(r));
        for (auto& cm : m->comms) cm = nullptr;
        trt_mgpu_destroy(m);
        return TRT_ERR_NCCL;
    }
    *out = m;
    return 0;
}

int trt_mgpu_destroy(trt_mgpu* m) {
    if (!m) return 0;
    DeviceGuard guard;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        cudaSetDevice(m->devices[g]);
        if (m->comms[g]) ncclCommDestroy(m->comms[g]);
        cudaFree(m->accum[g]);
        trt_destroy(m->ctx[g]);
    }
    delete m;
    return 0;
}

int trt_mgpu_count(const trt_mgpu* m) { return m ? (int)m->ctx.size() : 0; }

int trt_mgpu_upload_scene(trt_mgpu* m, const void* objects, int n_objects, const void* nodes, int n_nodes,
                          const int* lights, int n_lights, const trt_image* textures, int n_textures) {
    if (!m) return TRT_ERR_ARG;
    DeviceGuard guard;
    for (size_t g = 0; g < m->ctx.size(); g++)
        if (int rc = trt_upload_scene(m->ctx[g], objects, n_objects, nodes, n_nodes, lights, n_lights, textures, n_textures))
            return mfail("trt_upload_scene", trt_last_error()), rc;
    return 0;
}

int trt_mgpu_render_to_host(trt_mgpu* m, float* h_accum, int width, int height, int first_frame_seed, int n_frames,
                            const void* cam, const trt_opts* opts, float* pass_ms) {
    if (!m || !h_accum || !cam || width <= 0 || height <= 0 || n_frames < 0) return TRT_ERR_ARG;
    DeviceGuard guard;
    const int G = (int)m->ctx.size();
    const size_t bytes = (size_t)width * height * 16;
    if (bytes != m->accum_bytes) {
        m->accum_bytes = 0;  // stays 0 if an allocation below fails, so the next call allocates again
        for (int g = 0; g < G; g++) {
            MCU(cudaSetDevice(m->devices[g]));
            cudaFree(m->accum[g]);
            m->accum[g] = nullptr;
            MCU(cudaMalloc(&m->accum[g], bytes));
        }
        m->accum_bytes = bytes;
    }
    const auto t0 = std::chrono::steady_clock::now();
    // GPU g renders frames first+g, first+g+G, ...  (sample-index split; the union of the RNG streams
    // equals the single-GPU pass)
    std::vector<int> rc(G, 0);
    std::vector<std::string> err(G);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++) {
        workers.emplace_back([&, g]() {
            cudaSetDevice(m->devices[g]);
            trt_reset_counters(m->ctx[g]);
            cudaMemsetAsync(m->accum[g], 0, bytes, m->streams[g]);
            const int mine = n_frames > g ? (n_frames - g + G - 1) / G : 0;
            rc[g] = trt_render(m->ctx[g], m->accum[g], width, height, first_frame_seed + g, mine, G, cam, opts);
            if (rc[g]) err[g] = trt_last_error();
        });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < G; g++)
        if (rc[g]) return mfail("trt_render", err[g].c_str()), rc[g];
    // one all-reduce of the accumulation buffer per pass, in place, on each GPU's render stream
    MNCCL(ncclGroupStart());
    for (int g = 0; g < G; g++)
        MNCCL(ncclAllReduce(m->accum[g], m->accum[g], (size_t)width * height * 4, ncclFloat, ncclSum, m->comms[g], m->streams[g]));
    MNCCL(ncclGroupEnd());
    for (int g = 0; g < G; g++) {
        MCU(cudaSetDevice(m->devices[g]));
        MCU(cudaStreamSynchronize(m->streams[g]));
    }
    if (pass_ms) *pass_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    MCU(cudaSetDevice(m->devices[0]));
    MCU(cudaMemcpy(h_accum, m->accum[0], bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int trt_mgpu_render_accumulate(trt_mgpu* m, float* d_accum, int width, int height, int first_frame_seed, int n_frames,
                               const void* cam, const trt_opts* opts, float* pass_ms) {
    if (!m || !d_accum || !cam || width <= 0 || height <= 0 || n_frames < 0) return TRT_ERR_ARG;
    DeviceGuard guard;
    const int G = (int)m->ctx.size();
    const size_t bytes = (size_t)width * height * 16;
    if (bytes != m->accum_bytes) {
        m->accum_bytes = 0;
        for (int g = 0; g < G; g++) {
            MCU(cudaSetDevice(m->devices[g]));
            cudaFree(m->accum[g]);
            m->accum[g] = nullptr;
            MCU(cudaMalloc(&m->accum[g], bytes));
        }
        m->accum_bytes = bytes;
    }
    MCU(cudaSetDevice(m->devices[0]));
    MCU(cudaStreamSynchronize(cudaStreamLegacy));  // the caller's cudaMemset / snapshot copies of d_accum
    const auto t0 = std::chrono::steady_clock::now();
    const int active = n_frames < G ? (n_frames > 0 ? n_frames : 1) : G;  // GPUs that have a frame to render
    std::vector<int> rc(G, 0);
    std::vector<std::string> err(G);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++) {
        workers.emplace_back([&, g]() {
            cudaSetDevice(m->devices[g]);
            trt_reset_counters(m->ctx[g]);
            float* target = g == 0 ? d_accum : m->accum[g];
            if (g != 0 && active > 1) cudaMemsetAsync(target, 0, bytes, m->streams[g]);
            const int mine = n_frames > g ? (n_frames - g + G - 1) / G : 0;
            if (mine > 0) rc[g] = trt_render(m->ctx[g], target, width, height, first_frame_seed + g, mine, G, cam, opts);
            if (rc[g]) err[g] = trt_last_error();
        });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < G; g++)
        if (rc[g]) return mfail("trt_render", err[g].c_str()), rc[g];
    if (active > 1) {  // one reduction per pass, to the GPU that owns the caller's buffer (in place there)
        MNCCL(ncclGroupStart());
        for (int g = 0; g < G; g++) {
            float* buf = g == 0 ? d_accum : m->accum[g];
            MNCCL(ncclReduce(buf, buf, (size_t)width * height * 4, ncclFloat, ncclSum, 0, m->comms[g], m->streams[g]));
        }
        MNCCL(ncclGroupEnd());
    }
    for (int g = 0; g < G; g++) {
        MCU(cudaSetDevice(m->devices[g]));
        MCU(cudaStreamSynchronize(m->streams[g]));
    }
    MCU(cudaSetDevice(m->devices[0]));
    if (pass_ms) *pass_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

int trt_mgpu_rays(trt_mgpu* m, uint64_t* closest, uint64_t* shadow) {
    if (!m) return TRT_ERR_ARG;
    DeviceGuard guard;
    uint64_t c = 0, s = 0;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        trt_counters k;
        if (int rc = trt_get_counters(m->ctx[g], &k)) return rc;
        c += k.closest_rays;
        s += k.shadow_rays;
    }
    if (closest) *closest = c;
    if (shadow) *shadow = s;
    return 0;
}

}  // extern "C"
