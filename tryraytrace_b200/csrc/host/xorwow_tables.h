// xorwow_tables.h -- GF(2) skip-ahead matrices for the XORWOW generator.
//
// The reference seeds one cuRAND XORWOW stream per pixel per frame:
// curand_init(1984+frame, subsequence=pixel, 0) (reference src/renderer.cu:326).
// cuRAND's subsequence s starts s * 2^67 draws into the seed's sequence.  The v[0..4]
// part of the state advances linearly over GF(2) (Marsaglia xorshift, the step at
// curand_kernel.h:863-874), so "skip one subsequence" is a fixed 160x160 bit matrix
//     M = T^(2^67),   T = one-step transition,
// and the state of pixel p in frame f is  M^p * v0(f)  with v0(f) the seed scramble of
// curand_kernel.h:800-812.  cuRAND evaluates M^p digit by digit with ~16 mat-vecs per
// thread per launch; here the power is split by image row and column,
//     M^(row*w + col) = (M^w)^row * M^col,
// so a pixel costs ONE mat-vec with a row matrix that is uniform across a warp
// (kernels/xorwow.cuh), applied to a per-frame column vector table.
//
// Everything here is derived from the published algorithm; no cuRAND table is copied.
#pragma once
#include <cstdint>
#include <vector>

namespace trt {

// 160x160 bit matrix, column-major by input bit: col[b] is the image of basis vector e_b,
// b = 32*word + bit.  (Same orientation as cuRAND's precalc tables, curand_kernel.h:316-333.)
struct Gf2Mat {
    uint32_t col[160][5];
};

void gf2_identity(Gf2Mat& m);
void gf2_matvec(const Gf2Mat& m, const uint32_t v[5], uint32_t out[5]);
void gf2_matmul(const Gf2Mat& a, const Gf2Mat& b, Gf2Mat& out);  // out = a * b  (apply b first)
void gf2_square_n(Gf2Mat& m, int n);                              // m = m^(2^n)
void gf2_pow(const Gf2Mat& m, uint64_t e, Gf2Mat& out);

// T: one XORWOW draw.  M: one subsequence (2^67 draws).
void xorwow_step_matrix(Gf2Mat& t);
const Gf2Mat& xorwow_subsequence_matrix();

// Seed scramble of curand_init: v[0..4] and d for a 64-bit seed.
void xorwow_seed_state(uint64_t seed, uint32_t v[5], uint32_t* d);

// Full host evaluation (tests / small cases): state after curand_init(seed, subsequence, 0).
void xorwow_init_host(uint64_t seed, uint64_t subsequence, uint32_t v[5], uint32_t* d);

// Tables for an image of width w and height h:
//   row_mats[r]   = M^(w*r)              r in [0,h)       (h matrices of 800 words)
//   col_pows[j]   = M^(2^j)              j in [0,n_col_bits)  with 2^n_col_bits >= w
void xorwow_build_row_matrices(int w, int h, std::vector<Gf2Mat>& row_mats);
void xorwow_build_col_powers(int w, std::vector<Gf2Mat>& col_pows);
// Two-level column tables: lo[b] = M^b for b in [0,64), hi[a] = M^(64a) for a in [0, ceil(w/64)); M^col =
// hi[col >> 6] * lo[col & 63] (powers of one matrix commute), so the per-job column vectors cost two windowed
// mat-vecs per column instead of one full mat-vec per set bit of the column index.
void xorwow_build_col_levels(int w, std::vector<Gf2Mat>& lo, std::vector<Gf2Mat>& hi);

// 4-bit window form of a matrix for the device mat-vec (kernels/xorwow.cuh): the 160 input
// bits are cut into 40 nibbles; entry [n][v] is the XOR of the columns selected by the bits of
// v in nibble n, so a mat-vec is 40 table lookups instead of 160 conditional XORs.  Words 0-3
// go to `a` (16 B per entry), word 4 to `b`; 640 entries per matrix each.
constexpr int kXwWindowEntriesHost = 40 * 16;
void xorwow_window_table(const Gf2Mat& m, uint32_t* a /* 640 x 4 words */, uint32_t* b /* 640 words */);

}  // namespace trt
