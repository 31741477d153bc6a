// xorwow_tables.cpp -- see xorwow_tables.h.
#include "xorwow_tables.h"
#include <cstring>
#include <mutex>

namespace trt {

void gf2_identity(Gf2Mat& m) {
    std::memset(&m, 0, sizeof(m));
    for (int b = 0; b < 160; b++) m.col[b][b >> 5] = 1u << (b & 31);
}

void gf2_matvec(const Gf2Mat& m, const uint32_t v[5], uint32_t out[5]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    for (int w = 0; w < 5; w++) {
        uint32_t bits = v[w];
        while (bits) {
            const int j = __builtin_ctz(bits);
            bits &= bits - 1;
            const uint32_t* c = m.col[32 * w + j];
            r0 ^= c[0]; r1 ^= c[1]; r2 ^= c[2]; r3 ^= c[3]; r4 ^= c[4];
        }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

void gf2_matmul(const Gf2Mat& a, const Gf2Mat& b, Gf2Mat& out) {
    Gf2Mat tmp;
    for (int c = 0; c < 160; c++) gf2_matvec(a, b.col[c], tmp.col[c]);
    out = tmp;
}

void gf2_square_n(Gf2Mat& m, int n) {
    for (int i = 0; i < n; i++) gf2_matmul(m, m, m);
}

void gf2_pow(const Gf2Mat& m, uint64_t e, Gf2Mat& out) {
    Gf2Mat acc, base = m;
    gf2_identity(acc);
    while (e) {
        if (e & 1) gf2_matmul(base, acc, acc);
        e >>= 1;
        if (e) gf2_matmul(base, base, base);
    }
    out = acc;
}

void xorwow_step_matrix(Gf2Mat& t) {
    for (int b = 0; b < 160; b++) {
        uint32_t v[5] = {0, 0, 0, 0, 0};
        v[b >> 5] = 1u << (b & 31);
        // one draw, curand_kernel.h:863-872, on the v part only
        const uint32_t x = v[0] ^ (v[0] >> 2);
        const uint32_t n4 = (v[4] ^ (v[4] << 4)) ^ (x ^ (x << 1));
        t.col[b][0] = v[1];
        t.col[b][1] = v[2];
        t.col[b][2] = v[3];
        t.col[b][3] = v[4];
        t.col[b][4] = n4;
    }
}

const Gf2Mat& xorwow_subsequence_matrix() {
    static Gf2Mat m;
    static std::once_flag once;
    std::call_once(once, [] {
        xorwow_step_matrix(m);
        gf2_square_n(m, 67);  // 2^67 draws per subsequence
    });
    return m;
}

void xorwow_seed_state(uint64_t seed, uint32_t v[5], uint32_t* d) {
    const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    *d = 6615241u + t1 + t0;
    v[0] = 123456789u + t0;
    v[1] = 362436069u ^ t0;
    v[2] = 521288629u + t1;
    v[3] = 88675123u ^ t1;
    v[4] = 5783321u + t0;
}

void xorwow_init_host(uint64_t seed, uint64_t subsequence, uint32_t v[5], uint32_t* d) {
    uint32_t s[5];
    xorwow_seed_state(seed, s, d);
    Gf2Mat p;
    gf2_pow(xorwow_subsequence_matrix(), subsequence, p);
    gf2_matvec(p, s, v);
}

void xorwow_build_row_matrices(int w, int h, std::vector<Gf2Mat>& row_mats) {
    row_mats.resize(h);
    Gf2Mat step;
    gf2_pow(xorwow_subsequence_matrix(), (uint64_t)w, step);  // M^w
    gf2_identity(row_mats[0]);
    for (int r = 1; r < h; r++) gf2_matmul(step, row_mats[r - 1], row_mats[r]);
}

void xorwow_window_table(const Gf2Mat& m, uint32_t* a, uint32_t* b) {
    for (int n = 0; n < 40; n++) {
        for (int v = 0; v < 16; v++) {
            uint32_t acc[5] = {0, 0, 0, 0, 0};
            for (int bit = 0; bit < 4; bit++)
                if ((v >> bit) & 1)
                    for (int k = 0; k < 5; k++) acc[k] ^= m.col[4 * n + bit][k];
            const int e = n * 16 + v;
            for (int k = 0; k < 4; k++) a[e * 4 + k] = acc[k];
            b[e] = acc[4];
        }
    }
}

void xorwow_build_col_powers(int w, std::vector<Gf2Mat>& col_pows) {
    int bits = 1;
    while ((1 << bits) < w) bits++;
    col_pows.resize(bits);
    col_pows[0] = xorwow_subsequence_matrix();
    for (int j = 1; j < bits; j++) gf2_matmul(col_pows[j - 1], col_pows[j - 1], col_pows[j]);
}

void xorwow_build_col_levels(int w, std::vector<Gf2Mat>& lo, std::vector<Gf2Mat>& hi) {
    lo.resize(64);
    gf2_identity(lo[0]);
    for (int b = 1; b < 64; b++) gf2_matmul(xorwow_subsequence_matrix(), lo[b - 1], lo[b]);
    Gf2Mat step;
    gf2_matmul(xorwow_subsequence_matrix(), lo[63], step);  // M^64
    hi.resize((size_t)(w + 63) / 64);
    gf2_identity(hi[0]);
    for (size_t a = 1; a < hi.size(); a++) gf2_matmul(step, hi[a - 1], hi[a]);
}

}  // namespace trt
