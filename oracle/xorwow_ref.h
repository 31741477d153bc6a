// xorwow_ref.h -- TEST INFRASTRUCTURE (oracle).  CPU restatement of cuRAND's XORWOW
// generator as the reference uses it (/root/reference/src/renderer.cu:326, :331...):
//   curand_init(seed, subsequence, 0)   CUDA 12.9 curand_kernel.h:800-822
//   skipahead_sequence                  curand_kernel.h:721-736 (base-4 digits, matrix per digit)
//   curand()                            curand_kernel.h:863-874
//   curand_uniform()                    curand_uniform.h:69-72
// cuRAND (a toolkit header, not part of /root/reference) ships precomputed skip matrices;
// here they are derived from the published algorithm: the one-draw transition T of the
// xorshift part is linear over GF(2), one subsequence is 2^67 draws, so matrix k of the
// table is (T^(2^67))^(4^k).  Pinned by tests/test_xorwow.py against the golden states of
// SURVEY Appendix B.3 and against the toolkit header evaluated on the host.
#pragma once
#include <stdint.h>
#include <string.h>

namespace xwref {

struct State {
    uint32_t v[5];
    uint32_t d;
};

struct Mat {
    uint32_t m[160 * 5];  // cuRAND layout: row (32*i + j) holds the image of bit j of v[i]
};

inline void matvec(const Mat& a, uint32_t v[5]) {  // __curand_matvec_inplace, curand_kernel.h:316-333
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 32; j++)
            if (v[i] & (1u << j))
                for (int k = 0; k < 5; k++) r[k] ^= a.m[5 * (i * 32 + j) + k];
    memcpy(v, r, sizeof(r));
}

inline void matmul(const Mat& a, const Mat& b, Mat& out) {  // out = a after b
    Mat t;
    for (int row = 0; row < 160; row++) {
        uint32_t v[5];
        memcpy(v, &b.m[5 * row], sizeof(v));
        matvec(a, v);
        memcpy(&t.m[5 * row], v, sizeof(v));
    }
    out = t;
}

inline void step_v(uint32_t v[5]) {
    uint32_t t = (v[0] ^ (v[0] >> 2));
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

// table[k] = (T^(2^67))^(4^k), k = 0..15 covers subsequences below 2^32
inline const Mat* sequence_matrices() {
    static Mat table[16];
    static bool ready = false;
    if (!ready) {
        Mat t;
        for (int b = 0; b < 160; b++) {
            uint32_t v[5] = {0, 0, 0, 0, 0};
            v[b >> 5] = 1u << (b & 31);
            step_v(v);
            memcpy(&t.m[5 * b], v, sizeof(v));
        }
        for (int i = 0; i < 67; i++) matmul(t, t, t);
        table[0] = t;
        for (int k = 1; k < 16; k++) {
            matmul(table[k - 1], table[k - 1], table[k]);
            matmul(table[k], table[k], table[k]);
        }
        ready = true;
    }
    return table;
}

inline void init(uint64_t seed, uint64_t subsequence, State* s) {
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
    const Mat* tab = sequence_matrices();
    int k = 0;
    uint64_t x = subsequence;
    while (x) {
        for (unsigned t = 0; t < (x & 3); t++) matvec(tab[k], s->v);
        x >>= 2;
        k++;
    }
}

inline uint32_t next(State* s) {
    step_v(s->v);
    s->d += 362437u;
    return s->v[4] + s->d;
}

inline float uniform(State* s) { return next(s) * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f); }

}  // namespace xwref
