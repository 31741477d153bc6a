// curand_host.cu -- TEST INFRASTRUCTURE (oracle).  Known-answer source for the XORWOW
// tests: curand_init(seed, subsequence, 0) evaluated ON THE HOST by the CUDA toolkit's
// own header (curand_kernel.h:800-822 with the host tables of curand_precalc.h), so it
// runs without a GPU.  The header is host-callable once QUALIFIERS carries __host__.
#define QUALIFIERS static inline __host__ __device__
#include <curand_kernel.h>

extern "C" {

// out: 6 words per entry = v[0..4], d
int ref_xorwow_states(unsigned long long seed, const int* pixels, int n, unsigned* out) {
    for (int i = 0; i < n; i++) {
        curandState st;
        curand_init(seed, (unsigned long long)pixels[i], 0, &st);
        for (int k = 0; k < 5; k++) out[6 * i + k] = st.v[k];
        out[6 * i + 5] = st.d;
    }
    return 0;
}

// the first `n_draws` curand() outputs and curand_uniform() floats of one stream
int ref_xorwow_draws(unsigned long long seed, int pixel, int n_draws, unsigned* out_u32, float* out_f32) {
    curandState a, b;
    curand_init(seed, (unsigned long long)pixel, 0, &a);
    b = a;
    for (int i = 0; i < n_draws; i++) {
        if (out_u32) out_u32[i] = curand(&a);
        if (out_f32) out_f32[i] = curand_uniform(&b);
    }
    return 0;
}

}  // extern "C"
