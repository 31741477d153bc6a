// xorwow.cuh -- device side of the cuRAND-compatible XORWOW streams.
//
// Draw-for-draw identical to curand()/curand_uniform() on a state produced by
// curand_init(seed, subsequence, 0) (reference src/renderer.cu:326, :331 ...), but the
// state is produced with one warp-uniform GF(2) mat-vec per pixel instead of cuRAND's
// ~16 (see host/xorwow_tables.h for the algebra).
#pragma once
#include "common.cuh"

namespace trt {

struct Xorwow {
    uint32_t v0, v1, v2, v3, v4, d;
};

// one draw: curand_kernel.h:863-874
TRT_DEV uint32_t xw_next(Xorwow& s) {
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1;
    s.v1 = s.v2;
    s.v2 = s.v3;
    s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}

// curand_uniform: x * 2^-32 + 2^-33, one FMA (curand_uniform.h:69-72)
TRT_DEV float xw_uniform(Xorwow& s) {
    return p_fma(__uint2float_rn(xw_next(s)), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Device matrix layout: 160 columns of 8 words (5 used, 3 pad) so a column is two
// aligned vector loads; 5120 bytes per matrix.
constexpr int kXwMatWords = 160 * 8;

// out = mat * in over GF(2).  `mat` is the same address for every lane of a warp in the
// common case, so the loads are broadcasts; the conditional XOR is a mask-and-xor (LOP3).
TRT_DEV void xw_matvec(const uint32_t* __restrict__ mat, const uint32_t in[5], uint32_t out[5]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
    for (int w = 0; w < 5; w++) {
        const uint32_t bits = in[w];
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const uint4 c = __ldg(reinterpret_cast<const uint4*>(mat + (32 * w + j) * 8));
            const uint32_t c4 = __ldg(mat + (32 * w + j) * 8 + 4);
            const uint32_t m = 0u - ((bits >> j) & 1u);
            r0 ^= c.x & m;
            r1 ^= c.y & m;
            r2 ^= c.z & m;
            r3 ^= c.w & m;
            r4 ^= c4 & m;
        }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

// out = mat * in through the 4-bit window tables of host/xorwow_tables.h: 40 lookups.  The
// table of one image row is shared by every sample of that row, so the 20-byte entries a warp
// touches sit in a 12.8 KB L1-resident block; the 80 loads are independent of one another.
constexpr int kXwWindowEntries = 40 * 16;
TRT_DEV void xw_matvec_window(const uint4* __restrict__ ta, const uint32_t* __restrict__ tb, const uint32_t in[5],
                              uint32_t out[5]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
    for (int w = 0; w < 5; w++) {
        const uint32_t bits = in[w];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t e = (uint32_t)((w * 8 + k) * 16) + ((bits >> (4 * k)) & 15u);
            const uint4 a = __ldg(ta + e);
            const uint32_t b = __ldg(tb + e);
            r0 ^= a.x;
            r1 ^= a.y;
            r2 ^= a.z;
            r3 ^= a.w;
            r4 ^= b;
        }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

// seed scramble of curand_init (curand_kernel.h:800-812) for a seed that fits 32 bits
TRT_DEV void xw_seed(uint32_t seed_lo, uint32_t v[5], uint32_t* d) {
    const uint32_t s0 = seed_lo ^ 0xaad26b49u;
    const uint32_t s1 = 0xf7dcefddu;  // high seed word is zero
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    *d = 6615241u + t1 + t0;
    v[0] = 123456789u + t0;
    v[1] = 362436069u ^ t0;
    v[2] = 521288629u + t1;
    v[3] = 88675123u ^ t1;
    v[4] = 5783321u + t0;
}

// Column-vector table entry: M^col * v0(frame), plus the frame's Weyl word d.  32 bytes.
struct XwColVec {
    uint32_t v[5];
    uint32_t d;
    uint32_t pad[2];
};

}  // namespace trt
