// peaks.cu -- the two machine peaks the roofline of SURVEY 8(d) needs and MEASURED_PEAKS.json does not hold:
//   * FP32 FMA throughput at the sustained clock (dependent-free FFMA chains, all SMs, >= 100 ms), reported as
//     TFLOP/s (2 flop per FMA) and as T FP32-instructions/s, and the same through the packed FFMA2 form;
//   * L2-resident read bandwidth (a working set that fits the 126 MB L2, read repeatedly with 128-bit loads).
// Prints one JSON object.  Build: make peaks  ->  build/peaks.   Usage: build/peaks [device]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                   \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

constexpr int kChains = 8;      // independent accumulators per thread (covers the FMA latency)
constexpr int kInner = 4096;    // FMAs per chain per thread

__global__ void __launch_bounds__(1024) k_fma(float* out, float a, float b) {
    float acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; i++) acc[i] = (float)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < kInner / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < kChains; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; i++) s += acc[i];
    if (s == 12345.678f) out[0] = s;  // never true: keeps the chains alive
}

__device__ __forceinline__ float2 ffma2(float2 x, float2 y, float2 z) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.ftz.f32x2 rr, ra, rb, rc;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y), "f"(z.x), "f"(z.y));
    return r;
}

__global__ void __launch_bounds__(1024) k_fma2(float* out, float a, float b) {
    float2 acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; i++) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < kInner / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < kChains; i++) acc[i] = ffma2(acc[i], aa, bb);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; i++) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) out[0] = s;
}

// grid-stride sweeps over a buffer that fits the L2 (the warm-up launch and the first sweep bring it in; the
// following sweeps hit), 128-bit ld.global.cg loads (L1 bypassed), four loads in flight per thread
__global__ void __launch_bounds__(1024) k_l2_read(const uint4* __restrict__ buf, size_t n_vec, int sweeps, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int s = 0; s < sweeps; s++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride),
                        d = __ldcg(buf + i + 3 * stride);
            acc ^= a.x ^ b.y ^ c.z ^ d.w;
        }
        for (; i < n_vec; i += stride) acc ^= __ldcg(buf + i).x;
    }
    if (acc == 0x12345u) out[0] = acc;
}

int main(int argc, char** argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
    float* d_out;
    CK(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    const int blocks = p.multiProcessorCount * 2, threads = 1024;
    auto time_kernel = [&](auto launch, int reps) -> float {
        launch();  // warm-up
        cudaDeviceSynchronize();
        float best_total = 0.f;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&best_total, e0, e1);
        return best_total / reps;
    };
    // ~100+ ms of sustained work each
    const double fma_per_launch = (double)blocks * threads * kChains * kInner;
    const float ms1 = time_kernel([&] { k_fma<<<blocks, threads>>>(d_out, 1.0000001f, 1e-9f); }, 200);
    const float ms2 = time_kernel([&] { k_fma2<<<blocks, threads>>>(d_out, 1.0000001f, 1e-9f); }, 200);
    CK(cudaGetLastError());
    const double fp32_tflops = 2.0 * fma_per_launch / (ms1 * 1e-3) / 1e12;
    const double fp32x2_tflops = 4.0 * fma_per_launch / (ms2 * 1e-3) / 1e12;

    // L2-resident read: 48 MB working set (well inside the 126 MB L2, two dies)
    std::vector<double> l2;
    for (size_t mb : {24, 48, 96}) {
        const size_t bytes = mb << 20, n_vec = bytes / 16;
        uint4* buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        const int sweeps = 8;
        const float ms = time_kernel([&] { k_l2_read<<<blocks, threads>>>(buf, n_vec, sweeps, (unsigned*)d_out); }, 20);
        CK(cudaGetLastError());
        l2.push_back((double)bytes * sweeps / (ms * 1e-3) / 1e9);
        cudaFree(buf);
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f, "
           "\"fp32_fma_tflops\": %.2f, \"fp32_fma_tinstr_per_s\": %.2f, \"fp32x2_fma_tflops\": %.2f, "
           "\"fp32_nominal_tflops\": %.2f, "
           "\"l2_read_gbs_24mb\": %.0f, \"l2_read_gbs_48mb\": %.0f, \"l2_read_gbs_96mb\": %.0f, "
           "\"how\": \"FFMA: %d independent chains x %d per thread, %d x %d threads, 200 launches; "
           "L2: 8 grid-stride sweeps per launch with ld.global.cg.v4, 20 launches\"}\n",
           p.name, p.multiProcessorCount, clock_khz / 1e3, fp32_tflops, fp32_tflops / 2.0, fp32x2_tflops,
           p.multiProcessorCount * 128 * 2.0 * clock_khz * 1e3 / 1e12, l2[0], l2[1], l2[2], kChains, kInner, blocks, threads);
    return 0;
}
