// wavefront.cu -- kernels of the streaming wavefront path tracer (see wavefront.cuh).
//
// Replaces the reference megakernel render_kernel_impl (reference src/renderer.cu:317-760):
// regenerate = :319-356 (RNG seeding + primary ray), extend = :371-425, shade = :434-733 and
// :739-759, shadow = :273-314/:692-709.
#include "wavefront.cuh"
#include "raygen.cuh"
#include "shade.cuh"
#include "traverse_fast.cuh"
#include "traverse_ref.cuh"
#include "trt_capi.h"
#include <cstdlib>

namespace trt {

namespace {

constexpr int kBlock = 256;
constexpr int kShadeMaxBlock = 512;
constexpr int kCompactMinCap = 32 * 1024;  // the pool is not compacted below this many slots

TRT_DEV int pack_flags(int state, int depth, int mode) { return state | (depth << 8) | (mode << 16); }

// A paired 32-byte record (PoolView::od, ::rs) is LOADED with one 256-bit access (sm_100 LDG.E.256): a warp's 32
// records are 1 KB contiguous, so the load is fully coalesced, where two 128-bit loads at a 32-byte stride each use
// half of every sector they touch (shade: -4 %).  Records are STORED as two 128-bit halves: the halves of a sector
// written by one thread merge in L2; the 256-bit store form (st.global.v8) was measured equal (shade -0.8 %) and
// is not used.
TRT_DEV void ld_rec(const float4* rec, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(rec));
}
TRT_DEV void ld_rec(const uint4* rec, uint4& a, uint4& b) {
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(rec));
}
TRT_DEV void st_rec(float4* rec, const float4 a, const float4 b) { rec[0] = a; rec[1] = b; }
TRT_DEV void st_rec(uint4* rec, const uint4 a, const uint4 b) { rec[0] = a; rec[1] = b; }

__global__ void k_begin_job(Control* ctl, unsigned long long total, int capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->next_sample = 0;
    ctl->total_samples = total;
    ctl->n_free = 0;         // every slot is marked ended in the dead mask (k_init_pool): the first refill fills them all
    ctl->active_cap = capacity;
    ctl->compact_go = 0;
    ctl->alive = capacity;   // refill subtracts the slots that ended and adds the samples it starts
    ctl->n_regen = 0;
    ctl->regen_base = 0;
    ctl->refill_ticket = 0;
}

__global__ void k_reset_counters(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->cnt_samples = ctl->cnt_closest = ctl->cnt_shadow = ctl->cnt_nodes = ctl->cnt_tris = 0;
    ctl->cnt_replays = ctl->cnt_iterations = 0;
    ctl->cnt_nodes_closest = ctl->cnt_tris_closest = 0;
    ctl->cnt_tree_closest = ctl->cnt_tree_shadow = 0;
    for (int i = 0; i < 16; i++) ctl->dbg[i] = 0;
    ctl->cnt_violations = 0;
}

__global__ void k_reset_cursors(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->cursor_extend = 0;
    ctl->cursor_shadow = 0;
    ctl->active_cap = 0x7fffffff;  // scratch pools of the parity entry points are visited whole
}

// ---- drain-phase compaction -----------------------------------------------------------------
// When a job has handed out its last sample the pool thins out: paths end, nothing refills their
// slots, and every kernel keeps paying for the dead ones (a 32-slot chunk with three live rays
// costs the traversal kernels a full top phase).  Each time the live paths drop to 3/4 of the visited
// slots (LaunchDims::compact_quarters), the live
// slots beyond the new bound are moved into dead slots below it and active_cap shrinks, so the
// kernels of the remaining iterations visit a dense prefix of the pool.
// scan: A = live slots in [new_cap, active_cap), B = dead slots in [0, new_cap)   (|B| >= |A|)
__global__ void __launch_bounds__(1024) k_compact_scan(PoolView pool, Control* ctl, int* list_a, int* list_b) {
    if (!ctl->compact_go) return;
    // list positions: shared-memory atomics within the block, one global atomic per block, list and round
    // (a global atomic per warp is a quarter of a million adds on two addresses: 0.4 ms at 8 Mi slots)
    __shared__ int count[2], base[2];
    const int cap = ctl->active_cap, new_cap = ctl->compact_new_cap;
    const unsigned lane = threadIdx.x & 31u;
    for (int first = blockIdx.x * blockDim.x; first < cap; first += gridDim.x * blockDim.x) {
        if (threadIdx.x < 2) count[threadIdx.x] = 0;
        __syncthreads();
        const int slot = first + threadIdx.x;
        const bool live = slot < cap && (f2i(pool.od[2 * slot + 1].w) & 0xff) != SLOT_DEAD;
        const bool to_a = live && slot >= new_cap, to_b = !live && slot < new_cap;
        const unsigned ma = __ballot_sync(0xffffffffu, to_a), mb = __ballot_sync(0xffffffffu, to_b);
        int ba = 0, bb = 0;
        if (lane == 0) {
            if (ma) ba = atomicAdd(&count[0], __popc(ma));
            if (mb) bb = atomicAdd(&count[1], __popc(mb));
        }
        __syncthreads();
        if (threadIdx.x < 2) base[threadIdx.x] = count[threadIdx.x] ? atomicAdd(threadIdx.x ? &ctl->compact_b : &ctl->compact_a, count[threadIdx.x]) : 0;
        __syncthreads();
        ba = __shfl_sync(0xffffffffu, ba, 0) + base[0];
        bb = __shfl_sync(0xffffffffu, bb, 0) + base[1];
        if (to_a) list_a[ba + __popc(ma & ((1u << lane) - 1u))] = slot;
        if (to_b) list_b[bb + __popc(mb & ((1u << lane) - 1u))] = slot;
    }
}

// move: slot A[i] -> slot B[i], every array of the pool
__global__ void __launch_bounds__(kBlock) k_compact_move(PoolView pool, Control* __restrict__ ctl,
                                                         const int* __restrict__ list_a, const int* __restrict__ list_b, int debug) {
    if (!ctl->compact_go) return;
    const int n = ctl->compact_a;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int src = list_a[i], dst = list_b[i];
        if (debug) {  // the move owns neither slot by construction: source must hold a path, destination must be dead and inside the new bound
            const bool ok = (f2i(pool.od[2 * src + 1].w) & 0xff) != SLOT_DEAD && (f2i(pool.od[2 * dst + 1].w) & 0xff) == SLOT_DEAD &&
                            !(pool.od[2 * dst].w > 0.f) && dst < ctl->compact_new_cap && src >= ctl->compact_new_cap && src < ctl->active_cap;
            if (!ok) atomicAdd(&ctl->cnt_violations, 1ull);
        }
        // `hit` stays behind: it is rewritten by the traversal before shade reads it.  The shadow ray moves along:
        // the shadow rays of the last shade pass are traced AFTER this compaction (src is never visited again)
        float4 o, d;
        uint4 ra, rb;
        ld_rec(pool.od + 2 * (size_t)src, o, d);
        ld_rec(pool.rs + 2 * (size_t)src, ra, rb);
        const float4 t = pool.thr[src], pe = pool.pend[src], sh = pool.sh_d[src];
        st_rec(pool.od + 2 * (size_t)dst, o, d);
        st_rec(pool.rs + 2 * (size_t)dst, ra, rb);
        pool.thr[dst] = t;
        pool.pend[dst] = pe;
        pool.sh_d[dst] = sh;
    }
}

__global__ void k_compact_commit(Control* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || !ctl->compact_go) return;
    ctl->active_cap = ctl->compact_new_cap;
    ctl->compact_go = 0;
}

__global__ void k_init_pool(PoolView pool) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.capacity) return;
    st_rec(pool.od + 2 * (size_t)i, make_float4(0.f, 0.f, 0.f, 0.f),  // no shadow ray
           make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC))));
    if (pool.dead_mask && (i & 31) == 0) pool.dead_mask[i >> 5] = 0xffffffffu;
}

// ---- XORWOW column table: col_vecs[f*w + col] = M^col * v0(frame f) --------------------
// M^col = M^(64 * (col >> 6)) * M^(col & 63): two windowed mat-vecs (80 look-ups) through the two-level tables of
// host/xorwow_tables.h, instead of one full mat-vec (160 column loads) per set bit of the column index -- the
// kernel runs once per job in front of the first refill, 154 us -> ~20 us of a 4 ms one-frame call.
// Table layout: entry t of `col_a` / `col_b` = window table of lo[t] for t < 64, of hi[t - 64] after that.
__global__ void k_col_table(const uint4* __restrict__ col_a, const uint32_t* __restrict__ col_b, int w, int first_frame_seed,
                            int frame_stride, int seed_base, int n_frames, XwColVec* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_frames * w) return;
    const int f = (int)(i / w), col = (int)(i % w);
    uint32_t v[5], t[5], d;
    xw_seed((uint32_t)(seed_base + first_frame_seed + f * frame_stride), v, &d);
    const size_t lo = (size_t)(col & 63) * kXwWindowEntries, hi = (size_t)(64 + (col >> 6)) * kXwWindowEntries;
    xw_matvec_window(col_a + lo, col_b + lo, v, t);
    xw_matvec_window(col_a + hi, col_b + hi, t, v);
    XwColVec e;
#pragma unroll
    for (int k = 0; k < 5; k++) e.v[k] = v[k];
    e.d = d;
    e.pad[0] = e.pad[1] = 0;
    out[i] = e;
}

// RNG state of (job-local frame f, pixel) = M^(w*row) * col_vecs[f][col]
TRT_DEV Xorwow sample_rng(const JobParams& job, int f, int row, int col) {
    const XwColVec* cv = job.col_vecs + (size_t)f * job.rc.width + col;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(cv));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(cv) + 2);
    const uint32_t in[5] = {a.x, a.y, a.z, a.w, b.x};
    uint32_t o[5];
    xw_matvec_window(job.row_a + (size_t)row * kXwWindowEntries, job.row_b + (size_t)row * kXwWindowEntries, in, o);
    Xorwow s;
    s.v0 = o[0]; s.v1 = o[1]; s.v2 = o[2]; s.v3 = o[3]; s.v4 = o[4];
    s.d = b.y;
    return s;
}

// ---- refill: free scan + bookkeeping + regeneration in ONE kernel ------------------------------
// One thread per dead-mask word (32 slots), kRefillBlock words per block.  A block
//   1. scans its mask words (the slots that ended in the last shade pass) into a slot list in shared memory,
//   2. reserves that many camera samples with one atomicAdd on Control::next_sample (clamped to the job),
//   3. regenerates them, kRefillBlock samples per batch: the samples of a batch are consecutive pixels of one
//      frame, i.e. they sit in one image row (two at a row boundary), so the 4-bit-window table of M^(w*row)
//      (12.8 KB, xorwow.cuh) is staged in shared memory as five word planes -- for a fixed window the 16 entries of
//      a plane fall into 16 different banks, so the 200 look-ups of a sample are conflict-free LDS.32 instead
//      of 80 scattered global loads; lanes whose row is not staged (images narrower than a batch) use the
//      global table,
//   4. the block that finishes last closes the iteration's bookkeeping (a single-thread kernel in round 1): clamps next_sample,
//      counts the samples, resets the traversal cursor and decides about a drain-phase compaction.
// Which sample lands in which slot depends on the order the blocks reserve their ranges; the image does not
// (a sample's RNG stream is a function of its frame and pixel alone).
constexpr int kRefillBlock = 256;
constexpr int kRefillSlots = kRefillBlock * 32;

TRT_DEV void xw_matvec_planes(const uint32_t* __restrict__ tab, const uint32_t in[5], uint32_t out[5]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
    for (int w = 0; w < 5; w++) {
        const uint32_t bits = in[w];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t e = (uint32_t)((w * 8 + k) * 16) + ((bits >> (4 * k)) & 15u);
            r0 ^= tab[e];
            r1 ^= tab[kXwWindowEntries + e];
            r2 ^= tab[2 * kXwWindowEntries + e];
            r3 ^= tab[3 * kXwWindowEntries + e];
            r4 ^= tab[4 * kXwWindowEntries + e];
        }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

__global__ void __launch_bounds__(kRefillBlock, 4) k_refill(PoolView pool, Control* ctl, JobParams job, int compact_quarters,
                                                            int samples_left, int debug) {
    __shared__ uint32_t s_tab[2][5 * kXwWindowEntries];
    __shared__ uint16_t s_list[kRefillSlots];
    __shared__ int s_warp[kRefillBlock / 32];
    __shared__ int s_take;
    __shared__ unsigned long long s_base;
    const int cap = ctl->active_cap;
    const int n_words = (cap + 31) >> 5;
    const int w0 = blockIdx.x * kRefillBlock;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (w0 < n_words) {
        // 1. dead slots of this block -> s_list (positions relative to the block's first slot)
        const int wi = w0 + (int)threadIdx.x;
        uint32_t m = wi < n_words ? pool.dead_mask[wi] : 0u;
        const int c = __popc(m);
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < kRefillBlock / 32; i++) {
            const int v = s_warp[i];
            if (i < (int)warp) before += v;
            total += v;
        }
        // 2. reserve the samples
        if (threadIdx.x == 0) {
            int take = 0;
            unsigned long long base = 0;
            if (total > 0 && samples_left) {
                base = atomicAdd(&ctl->next_sample, (unsigned long long)total);
                const unsigned long long all = ctl->total_samples;
                take = base >= all ? 0 : (int)min((unsigned long long)total, all - base);
            }
            if (total != take) atomicAdd(&ctl->alive, take - total);
            s_take = take;
            s_base = base;
        }
        if (samples_left) {
            int at = before + incl - c;
            const int rel = (int)threadIdx.x << 5;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                s_list[at++] = (uint16_t)(rel + b);
            }
        }
        __syncthreads();
        // 3. regenerate
        const int take = s_take;
        if (take > 0) {
            const unsigned long long base = s_base;
            const unsigned pixels = (unsigned)(job.rc.width * job.rc.height);
            const int first_slot = w0 << 5;
            int staged0 = -1, staged1 = -1;  // rows whose tables sit in s_tab[0] / s_tab[1] (block-uniform)
            for (int q = 0; q < take; q += kRefillBlock) {
                // frame / pixel of the batch's first sample (one 64-bit division per batch, not per sample) and the
                // rows of its first and last sample
                const unsigned long long sa = base + (unsigned long long)q;
                const unsigned long long fa = sa / pixels;
                const unsigned pa = (unsigned)(sa - fa * pixels);
                unsigned pb = pa + (unsigned)(min(kRefillBlock, take - q) - 1);
                while (pb >= pixels) pb -= pixels;
                const int row_a = (int)(pa / (unsigned)job.rc.width), row_b = (int)(pb / (unsigned)job.rc.width);
                const bool have_a = row_a == staged0 || row_a == staged1;
                const bool have_b = row_b == staged0 || row_b == staged1;
                if (!have_a || !have_b) {
                    __syncthreads();  // the previous batch has finished reading the tables
                    // row_a goes to the buffer that does not hold row_b (and vice versa)
                    int slot_a = -1, slot_b = -1;
                    if (!have_a) slot_a = (have_b && row_b == staged0) ? 1 : 0;
                    if (!have_b && row_b != row_a) slot_b = !have_a ? (slot_a ^ 1) : (row_a == staged0 ? 1 : 0);
                    for (int i = threadIdx.x; i < kXwWindowEntries; i += kRefillBlock) {
                        if (slot_a >= 0) {
                            const uint4 a = __ldg(job.row_a + (size_t)row_a * kXwWindowEntries + i);
                            const uint32_t b = __ldg(job.row_b + (size_t)row_a * kXwWindowEntries + i);
                            uint32_t* t = s_tab[slot_a];
                            t[i] = a.x; t[kXwWindowEntries + i] = a.y; t[2 * kXwWindowEntries + i] = a.z;
                            t[3 * kXwWindowEntries + i] = a.w; t[4 * kXwWindowEntries + i] = b;
                        }
                        if (slot_b >= 0) {
                            const uint4 a = __ldg(job.row_a + (size_t)row_b * kXwWindowEntries + i);
                            const uint32_t b = __ldg(job.row_b + (size_t)row_b * kXwWindowEntries + i);
                            uint32_t* t = s_tab[slot_b];
                            t[i] = a.x; t[kXwWindowEntries + i] = a.y; t[2 * kXwWindowEntries + i] = a.z;
                            t[3 * kXwWindowEntries + i] = a.w; t[4 * kXwWindowEntries + i] = b;
                        }
                    }
                    if (slot_a == 0) staged0 = row_a;
                    if (slot_a == 1) staged1 = row_a;
                    if (slot_b == 0) staged0 = row_b;
                    if (slot_b == 1) staged1 = row_b;
                    __syncthreads();
                }
                const int r = q + (int)threadIdx.x;
                if (r < take) {
                    const int slot = first_slot + (int)s_list[r];
                    unsigned upix = pa + threadIdx.x;
                    int f = (int)fa;
                    while (upix >= pixels) { upix -= pixels; f++; }
                    const int pix = (int)upix;                  // reference pixel index i
                    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
                    const int y = job.rc.height - 1 - row;      // i = (h-1-y)*w + x  (reference :322)
                    const XwColVec* cv = job.col_vecs + (size_t)f * job.rc.width + col;
                    const uint4 ca = __ldg(reinterpret_cast<const uint4*>(cv));
                    const uint2 cb = __ldg(reinterpret_cast<const uint2*>(cv) + 2);
                    const uint32_t in[5] = {ca.x, ca.y, ca.z, ca.w, cb.x};
                    uint32_t o[5];
                    if (row == staged0) xw_matvec_planes(s_tab[0], in, o);
                    else if (row == staged1) xw_matvec_planes(s_tab[1], in, o);
                    else xw_matvec_window(job.row_a + (size_t)row * kXwWindowEntries, job.row_b + (size_t)row * kXwWindowEntries, in, o);
                    Xorwow rng;
                    rng.v0 = o[0]; rng.v1 = o[1]; rng.v2 = o[2]; rng.v3 = o[3]; rng.v4 = o[4];
                    rng.d = cb.y;
                    const Ray ray = primary_ray(job.cam, col, y, job.rc.width, job.rc.height, rng);
                    if (debug) {  // the slot must have ended (dead, no shadow ray waiting) and lie inside the visited prefix
                        const bool ok = (f2i(pool.od[2 * slot + 1].w) & 0xff) == SLOT_DEAD && !(pool.od[2 * slot].w > 0.f) && slot < cap;
                        if (!ok) atomicAdd(&ctl->cnt_violations, 1ull);
                    }
                    // Two full 32-byte sectors per fresh path (a 16-byte store into a sector nobody else is writing
                    // costs a read-modify-write of the sector).  A fresh path has throughput 1 and no radiance:
                    // thr / pend are not written; the pixel index rides in od[2s].w with the sign bit set (that word
                    // carries a shadow-ray length from depth 1 on, and "> 0" means "trace it") until shade moves it
                    // to thr.w
                    st_rec(pool.od + 2 * (size_t)slot, make_float4(ray.o.x, ray.o.y, ray.o.z, i2f(pix | kFreshPixelBit)),
                           make_float4(ray.d.x, ray.d.y, ray.d.z, i2f(pack_flags(SLOT_ACTIVE, 0, MODE_SPEC))));
                    st_rec(pool.rs + 2 * (size_t)slot, make_uint4(rng.v0, rng.v1, rng.v2, rng.v3), make_uint4(rng.v4, rng.d, 0u, 0u));
                }
            }
        }
    }
    // 4. the last block to get here closes the iteration's bookkeeping
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&ctl->refill_ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long all = ctl->total_samples;
        unsigned long long next = *(volatile unsigned long long*)&ctl->next_sample;
        if (next > all) next = all;
        const unsigned long long started = next - ctl->regen_base;  // regen_base: next_sample after the previous refill
        ctl->next_sample = next;
        ctl->regen_base = next;
        ctl->n_regen = (int)started;
        ctl->cnt_samples += started;
        ctl->cnt_iterations += 1;
        ctl->cursor_extend = 0;  // cursor_shadow is reset by the shade kernel
        ctl->refill_ticket = 0;
        // drain phase: no sample left to start, and the live paths have dropped to the given share of the visited slots
        const int alive = *(volatile int*)&ctl->alive;
        const bool go = started == 0 && next == all && cap > kCompactMinCap && (long long)alive * 4 <= (long long)cap * compact_quarters;
        ctl->compact_go = go ? 1 : 0;
        if (go) {
            ctl->compact_new_cap = max(kCompactMinCap, (alive + kShadeMaxBlock - 1) / kShadeMaxBlock * kShadeMaxBlock);
            ctl->compact_a = ctl->compact_b = 0;
        }
    }
}

// warp-aggregated add of per-thread counters (one atomic per warp)
TRT_DEV void warp_add(unsigned long long* counter, unsigned v) {
    const unsigned total = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0 && total) atomicAdd(counter, (unsigned long long)total);
}

// an occluded shadow ray cancels the next-event contribution waiting in pend.xyz (pend.w is radiance.z: it stays)
TRT_DEV void cancel_pend(const PoolView& pool, int slot) {
    float* p = reinterpret_cast<float*>(pool.pend + slot);
    __stcs(reinterpret_cast<float2*>(p), make_float2(0.f, 0.f));
    __stcs(p + 2, 0.f);
}

// ---- extend, reference order: closest hit for every active slot --------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_extend_ref(PoolView pool, SceneDev sc, Control* ctl) {
    unsigned nodes = 0, tris = 0, rays = 0;
    const int cap = min(pool.capacity, ctl->active_cap);
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < cap; slot += gridDim.x * blockDim.x) {
        const float4 d4 = pool.od[2 * slot + 1];
        if ((f2i(d4.w) & 0xff) != SLOT_ACTIVE) continue;
        const float4 o4 = pool.od[2 * slot];
        Ray r;
        r.o = f3(o4.x, o4.y, o4.z);
        r.d = f3(d4.x, d4.y, d4.z);
        float t;
        VisitCounts vc = {0, 0, 0};
        const int id = ref_closest<COUNT>(sc, r, &t, &vc);
        if (COUNT) { nodes += vc.fetched; tris += vc.tris; }
        rays++;
        pool.hit[slot] = make_float2(t, i2f(id));
    }
    warp_add(&ctl->cnt_closest, rays);
    if (COUNT) {
        warp_add(&ctl->cnt_nodes, nodes);
        warp_add(&ctl->cnt_tris, tris);
        warp_add(&ctl->cnt_nodes_closest, nodes);
        warp_add(&ctl->cnt_tris_closest, tris);
    }
}

// ---- shade: one thread per slot --------------------------------------------------------
// The kernel is a stream over ~250 bytes of path state per slot and is bound by how many of those
// loads are in flight, so while the pool is mostly live (`eager`) every array of the slot is
// requested up front, before the state word has come back; in the drain phase of a job, when
// most slots are dead, the loads wait for the state check instead.
template <bool COUNT, bool FAST, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_shade(PoolView pool, Control* ctl, SceneDev sc, JobParams job, int eager) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;  // capacity is a multiple of the block size
    if (slot == 0) ctl->cursor_shadow = 0;                   // the shadow kernel of this iteration starts at chunk 0
    // The state loads go out BEFORE anything that depends on the control block has come back (the grid is
    // sized by the host's bound, every slot below it is valid memory): a dependent read of the control block
    // first cost every CTA an L2 round trip before its first state load.  `eager` is the host's knowledge
    // (samples are still being handed out, the pool is mostly live).
    float4 o4, d4;
    ld_rec(pool.od + 2 * (size_t)slot, o4, d4);
    float4 thr4, pend4;
    float2 hit;
    uint4 ra, rb;
    if (eager) {
        thr4 = pool.thr[slot]; pend4 = pool.pend[slot];
        hit = pool.hit[slot]; ld_rec(pool.rs + 2 * (size_t)slot, ra, rb);
    }
    const int cap = ctl->active_cap;                         // a multiple of the block size
    if (blockIdx.x * blockDim.x >= cap) return;              // whole block beyond the visited prefix
    const int flags = f2i(d4.w);
    const int state = flags & 0xff;
    bool terminated = false;
    if (state != SLOT_DEAD) {
        if (!eager) {
            thr4 = pool.thr[slot]; pend4 = pool.pend[slot];
            hit = pool.hit[slot]; ld_rec(pool.rs + 2 * (size_t)slot, ra, rb);
        }
        PathVertexIO io;
        io.shadow = false;
        io.depth = (flags >> 8) & 0xff;
        const bool fresh = io.depth == 0;  // refill leaves thr / pend unwritten and the pixel in od[2s].w
        const int pix = fresh ? (f2i(o4.w) & ~kFreshPixelBit) : f2i(thr4.w);
        io.thr = fresh ? f3(1.f, 1.f, 1.f) : f3(thr4.x, thr4.y, thr4.z);
        // radiance: .x .y ride behind the RNG state, .z behind the pending contribution
        io.rad = fresh ? f3(0.f, 0.f, 0.f) : f3(__uint_as_float(rb.z), __uint_as_float(rb.w), pend4.w);
        // next-event estimate of the previous vertex; the shadow pass zeroed it if occluded
        if (io.depth > 0) io.rad = v_add(io.rad, f3(pend4.x, pend4.y, pend4.z));
        if (state == SLOT_ACTIVE) {
            float t_hit = hit.x;
            int id = f2i(hit.y);
            // FAST traversal leaves the winner unverified (traverse_fast.cuh, point 5); REF ids pass through
            if (FAST && resolve_hit(sc, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), t_hit, id))
                atomicAdd(&ctl->cnt_replays, 1ull);
            if (id < 0) {
                terminated = true;  // miss: black environment (reference :427)
            } else {
                io.ray.o = f3(o4.x, o4.y, o4.z);
                io.ray.d = f3(d4.x, d4.y, d4.z);
                io.prev_mode = (flags >> 16) & 0xff;
                io.rng.v0 = ra.x; io.rng.v1 = ra.y; io.rng.v2 = ra.z; io.rng.v3 = ra.w;
                io.rng.v4 = rb.x; io.rng.d = rb.y;
                if (!shade_vertex(sc, job.rc, io, id, t_hit)) {
                    terminated = true;
                } else {
                    const int depth = io.depth + 1;
                    // the reference loop ends after max_depth vertices; a path that still has a
                    // shadow ray in flight is finalised one iteration later (SLOT_FINISH)
                    const int ns = depth >= job.rc.max_depth ? SLOT_FINISH : SLOT_ACTIVE;
                    // the shadow ray starts where the next ray starts (x_hit + nl * 1e-3, reference
                    // :692 and :731): its length rides in od[2s].w (0 = none), its direction in sh_d
                    st_rec(pool.od + 2 * (size_t)slot, make_float4(io.ray.o.x, io.ray.o.y, io.ray.o.z, io.shadow ? io.shadow_max_dist : 0.f),
                           make_float4(io.ray.d.x, io.ray.d.y, io.ray.d.z, i2f(pack_flags(ns, depth, io.prev_mode))));
                    pool.thr[slot] = make_float4(io.thr.x, io.thr.y, io.thr.z, i2f(pix));
                    st_rec(pool.rs + 2 * (size_t)slot, make_uint4(io.rng.v0, io.rng.v1, io.rng.v2, io.rng.v3),
                           make_uint4(io.rng.v4, io.rng.d, __float_as_uint(io.rad.x), __float_as_uint(io.rad.y)));
                    if (io.shadow) {
                        pool.pend[slot] = make_float4(io.shadow_contrib.x, io.shadow_contrib.y, io.shadow_contrib.z, io.rad.z);
                        pool.sh_d[slot] = make_float4(io.shadow_ray.d.x, io.shadow_ray.d.y, io.shadow_ray.d.z, 0.f);
                    } else {
                        pool.pend[slot] = make_float4(0.f, 0.f, 0.f, io.rad.z);
                    }
                }
            }
        } else {
            terminated = true;  // SLOT_FINISH
        }
        if (terminated) {
            F3 rad = io.rad;
            if (filter_sample(rad)) {  // reference :739-759
                float* a = job.accum + (size_t)pix * 4;
                atomicAdd(a + 0, rad.x);
                atomicAdd(a + 1, rad.y);
                atomicAdd(a + 2, rad.z);
            }
            // one full sector: no shadow ray, slot dead
            st_rec(pool.od + 2 * (size_t)slot, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC))));
        }
    }
    // which slots ended: one mask word per warp (all 32 lanes are here); k_refill turns them into fresh paths
    const unsigned ended = __ballot_sync(0xffffffffu, terminated);
    if ((threadIdx.x & 31u) == 0) pool.dead_mask[slot >> 5] = ended;
}

// ---- shadow, reference order: any hit for every slot that holds a shadow ray ---------------
template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_shadow_ref(PoolView pool, SceneDev sc, Control* ctl) {
    unsigned nodes = 0, tris = 0, rays = 0;
    const int cap = min(pool.capacity, ctl->active_cap);
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < cap; slot += gridDim.x * blockDim.x) {
        const float4 o4 = pool.od[2 * slot];
        if (!(o4.w > 0.f)) continue;  // no shadow ray in this slot
        const float4 d4 = pool.sh_d[slot];
        Ray r;
        r.o = f3(o4.x, o4.y, o4.z);
        r.d = f3(d4.x, d4.y, d4.z);
        VisitCounts vc = {0, 0, 0};
        const bool occluded = ref_shadow<COUNT>(sc, r, o4.w, &vc);
        if (COUNT) { nodes += vc.fetched; tris += vc.tris; }
        rays++;
        if (occluded) cancel_pend(pool, slot);
    }
    warp_add(&ctl->cnt_shadow, rays);
    if (COUNT) {
        warp_add(&ctl->cnt_nodes, nodes);
        warp_add(&ctl->cnt_tris, tris);
    }
}

// ---- persistent fast-path kernels: two-level while-while traversal with dynamic ray fetch ------
// One CTA per SM.  Shared memory, in order:
//   [ top-of-tree nodes: k_smem x 128 B ][ per-thread stacks: S x THREADS entries ]
//   [ per-warp ray staging: 2 buffers x (32 origins + 32 directions) x 16 B ][ per-warp mbarriers ]
//   [ per-warp tree-ray queue: kQueueCap entries ]
// A warp repeatedly
//   1. takes a chunk of 32 consecutive pool slots -- lane 0 claims it with one atomicAdd on a global
//      cursor and issues two TMA bulk copies (origins, directions) that complete on an mbarrier;
//      the copy for chunk r+1 is in flight while chunk r is processed -- and runs the TOP PHASE on
//      it, one ray per lane, fully converged (traverse_fast.cuh).  Rays decided there write their
//      result at once; rays that enter the tree are compacted into the warp's queue;
//   2. refills idle traversal lanes from the queue when fewer than `refill_below` lanes hold a ray;
//   3. runs one traversal ROUND for the rays it holds (at most one node step and one triangle
//      step per lane).
constexpr int kChunk = 32;
constexpr int kStageBytesPerWarp = 2 * 2 * kChunk * 16;  // 2 buffers x (o + d) x 32 x float4
constexpr int kQueueCap = 48;                             // entries; the top phase runs while <= kQueueLow are queued
constexpr int kQueueLow = kQueueCap - kChunk;
constexpr int kQueueBytesClosest = kQueueCap * (16 + 16 + 4);
constexpr int kTopStageBytes = 3 * kMaxTop * 16;          // closest-hit kernel: shared-memory copy of the root-level list
constexpr int kQueueBytesShadow = kQueueCap * (16 + 16);

template <int THREADS> struct FastCfg;
template <> struct FastCfg<512>  { static constexpr int SC = 16, SS = 16; };
template <> struct FastCfg<768>  { static constexpr int SC = 12, SS = 12; };
template <> struct FastCfg<896>  { static constexpr int SC = 11, SS = 11; };
template <> struct FastCfg<1024> { static constexpr int SC = 8,  SS = 10; };

constexpr int kClaimChunks = 4;  // chunks a warp claims per atomic on the global cursor (one round trip per 128 slots)

struct Feeder {  // warp-uniform state of the staging double buffer
    int cur_buf, cur_base, nxt_base;
    int claim_base, claim_end;  // slots of the current claim not yet requested: [claim_base, claim_end)
    bool fresh;    // the current buffer holds a chunk that has not been processed yet
    bool nxt_req;  // a bulk copy into the other buffer has been issued
    bool drained;  // the global cursor ran past the pool
    unsigned phase;  // bit b = parity the next wait on buffer b uses
    float4 cur_sh, nxt_sh;  // any-hit work: this lane's shadow direction of the current / the requested chunk
};

TRT_DEV void feeder_init(Feeder& f) {
    f.cur_buf = 0;
    f.cur_base = 0;
    f.nxt_base = 0;
    f.claim_base = f.claim_end = 0;
    f.fresh = false;
    f.nxt_req = false;
    f.drained = false;
    f.phase = 0;
    f.cur_sh = f.nxt_sh = make_float4(0.f, 0.f, 0.f, 0.f);
}
TRT_DEV bool feeder_exhausted(const Feeder& f) { return f.drained && !f.fresh && !f.nxt_req; }

// Make the next chunk current if it has landed (or wait for it when `block`: nobody in the warp
// has anything else to do), and keep one chunk in flight behind it.  A chunk is the 32 origin/direction
// records of 32 consecutive slots: ONE bulk copy of 1 KB.  SHADOW: the any-hit work reads the origin (and the
// shadow length in its .w) from the same records and the shadow direction from sh_d; that vector is fetched by
// the lanes themselves (a coalesced 512-byte load) when the chunk is requested and waits in a register, so the
// staging buffers stay at 1 KB per chunk.
template <bool SHADOW>
TRT_DEV void feeder_advance(Feeder& f, float4* stage, uint64_t* bars, const float4* src_od, const float4* src_sh,
                            int* cursor, int limit, unsigned lane, bool block) {
    if (!f.fresh && f.nxt_req) {
        const int nb = f.cur_buf ^ 1;
        const uint32_t parity = (f.phase >> nb) & 1u;
        // every lane polls (one converged instruction, no branch around a single-lane section); the vote makes the
        // answer warp-uniform should the phase flip between two lanes' reads
        bool ready = mbar_try_wait(&bars[nb], parity);
        if (block) {
            while (!__all_sync(0xffffffffu, ready)) ready = mbar_try_wait(&bars[nb], parity);
        } else {
            ready = __all_sync(0xffffffffu, ready);
        }
        if (ready) {
            f.phase ^= 1u << nb;
            f.cur_buf = nb;
            f.cur_base = f.nxt_base;
            f.fresh = true;
            f.nxt_req = false;
            if (SHADOW) f.cur_sh = f.nxt_sh;
        }
    }
    if (!f.nxt_req && !f.drained) {
        if (f.claim_base >= f.claim_end) {  // claim the next kClaimChunks chunks
            int c = 0;
            if (lane == 0) c = atomicAdd(cursor, kClaimChunks * kChunk);
            c = __shfl_sync(0xffffffffu, c, 0);
            f.claim_base = c;
            f.claim_end = min(c + kClaimChunks * kChunk, limit);  // limit is a multiple of the chunk size
        }
        const int base = f.claim_base;
        f.claim_base += kChunk;
        if (base >= limit) {
            f.drained = true;
        } else {
            const int nb = f.cur_buf ^ 1;  // the buffer not being read
            if (lane == 0) {
                mbar_expect_tx(&bars[nb], 2 * kChunk * 16);
                bulk_g2s(stage + nb * (2 * kChunk), src_od + 2 * (size_t)base, 2 * kChunk * 16, &bars[nb]);
            }
            if (SHADOW) f.nxt_sh = __ldg(src_sh + base + lane);
            f.nxt_req = true;
            f.nxt_base = base;
        }
    }
}

// The closest-hit work of one warp: `stage` / `bars` / `queue` are the warp's own staging buffers, mbarriers and
// tree-ray queue, `stk_base` the shared-window address of this thread's first stack entry.
template <int THREADS, int S, bool COUNT, bool WIDE>
TRT_DEV void extend_phase(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl, int k_smem,
                          int refill_below, const Phases& ph, const unsigned char* s_nodes, uint32_t stk_base,
                          float4* stage, uint64_t* bars, unsigned char* queue, const float4* s_top) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    float4* q_o = reinterpret_cast<float4*>(queue);  // origin, d_min
    float4* q_d = q_o + kQueueCap;                    // direction, winner so far (id | flag, -1 none)
    int* q_s = reinterpret_cast<int*>(q_d + kQueueCap);  // pool slot
    constexpr uint32_t E = THREADS * 8;  // bytes between consecutive stack entries of one lane
    const uint32_t stk_ttop = stk_base + (S - 1) * E;
    uint2 spill[kSpillEntries];
    const int slot_limit = min(pool.capacity, ctl->active_cap);
    Feeder fd;
    feeder_init(fd);
    ClosestRay st;
    st.np = stk_base;
    st.tp = stk_ttop;
    st.nspill = 0;
    st.cur = kWideEmptyRef;
    WideCounts wc = {0, 0};
    unsigned wc_tree = 0;
    unsigned dbg_rounds = 0, dbg_has = 0, dbg_nsteps = 0, dbg_tsteps = 0, dbg_chunks = 0;
    bool has = false;
    int slot = -1;
    int qn = 0;  // warp-uniform: entries in the tree-ray queue
    unsigned rays = 0;
    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, has);
        // 1. top phase on fresh chunks while the queue has room for a whole chunk
        while (qn <= kQueueLow && !feeder_exhausted(fd)) {
            feeder_advance<false>(fd, stage, bars, pool.od, nullptr, &ctl->cursor_extend, slot_limit, lane,
                                  act == 0 && qn == 0);
            if (!fd.fresh) break;  // the next chunk has not landed: traverse meanwhile
            const float4* buf = stage + fd.cur_buf * (2 * kChunk);
            const float4 o4 = buf[2 * lane], d4 = buf[2 * lane + 1];
            const bool live = (f2i(d4.w) & 0xff) == SLOT_ACTIVE;
            const int my_slot = fd.cur_base + (int)lane;
            TopResult tr = ph.ranked_top ? top_closest_ranked(top, s_top, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), live)
                                         : top_closest(top, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
            if (live) rays++;
            // decided by the root-level list alone (the winner is verified by the consumer, resolve_hit)
            if (live && !tr.enters) st_cs_f2(&pool.hit[my_slot], make_float2(tr.d_min, i2f(tr.id)));
            const bool tree = live && tr.enters;
            const unsigned m = __ballot_sync(0xffffffffu, tree);
            if (tree) {
                const int qi = qn + __popc(m & lt_mask);
                q_o[qi] = make_float4(o4.x, o4.y, o4.z, tr.d_min);
                q_d[qi] = make_float4(d4.x, d4.y, d4.z, i2f(tr.id));
                q_s[qi] = my_slot;
            }
            qn += __popc(m);
            if (COUNT && lane == 0) { wc_tree += __popc(m); dbg_chunks++; }
            fd.fresh = false;
            __syncwarp();
        }
        // 2. refill idle lanes from the queue
        if (qn > 0 && __popc(act) < refill_below) {
            const unsigned idle = ~act;
            const int rank = __popc(idle & lt_mask);
            if (!has && rank < qn) {
                const int qi = qn - 1 - rank;
                const float4 o4 = q_o[qi], d4 = q_d[qi];
                closest_begin(st, o4, d4, o4.w, f2i(d4.w), stk_base, stk_ttop);
                slot = q_s[qi];
                has = true;
            }
            qn = max(0, qn - __popc(idle));
            act = __ballot_sync(0xffffffffu, has);
            __syncwarp();
        }
        if (act == 0) {
            if (qn == 0 && feeder_exhausted(fd)) break;
            continue;
        }
        if (COUNT && lane == 0) { dbg_rounds++; dbg_has += __popc(act); }
        // 3. node phase: up to `node_iters` node steps, while enough lanes still have node work
#pragma unroll 1
        for (int it = 0; it < ph.node_iters; it++) {
            if (COUNT && lane == 0) dbg_nsteps++;
            closest_node_step<E, S, COUNT, WIDE>(s_nodes, k_smem, sc, st, stk_base, spill, &wc);
            const unsigned m = __ballot_sync(0xffffffffu, closest_node_work(st, stk_base) && closest_has_room<E, S>(st));
            if (__popc(m) < ph.node_min) break;
        }
        // 4. triangle phase: at least one step, more while enough lanes have a triangle waiting
#pragma unroll 1
        for (;;) {
            if (COUNT && lane == 0) dbg_tsteps++;
            closest_tri_step2<E, S, COUNT>(sc, st, stk_base, &wc);
            const unsigned m = __ballot_sync(0xffffffffu, st.tp != stk_ttop);
            if (__popc(m) < ph.tri_min) break;
        }
        // 5. finished rays write their hit (verified by the consumer, resolve_hit)
        if (has && closest_done<E, S>(st, stk_base)) {
            st_cs_f2(&pool.hit[slot], make_float2(st.d_min, i2f(st.id)));
            has = false;
        }
    }
    warp_add(&ctl->cnt_closest, rays);
    if (COUNT) {
        warp_add(&ctl->cnt_nodes, wc.nodes);
        warp_add(&ctl->cnt_tris, wc.tris);
        warp_add(&ctl->cnt_nodes_closest, wc.nodes);
        warp_add(&ctl->cnt_tris_closest, wc.tris);
        warp_add(&ctl->cnt_tree_closest, wc_tree);
        warp_add(&ctl->dbg[0], dbg_rounds);
        warp_add(&ctl->dbg[1], dbg_has);
        warp_add(&ctl->dbg[2], dbg_nsteps);
        warp_add(&ctl->dbg[3], dbg_tsteps);
        warp_add(&ctl->dbg[4], dbg_chunks);
    }
}

// The any-hit work of one warp.  E = bytes between consecutive stack entries of a lane: THREADS * 4 in the stand-alone
// kernel (conflict free); THREADS * 8 in the combined kernel, where a thread's entries sit where its closest-hit
// entries sit, so that a warp can move on to the closest-hit rays while other warps are still tracing shadow rays.
template <int THREADS, int S, uint32_t E, bool COUNT, bool WIDE, bool PAIR>
TRT_DEV void shadow_phase(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl, int k_smem,
                          int refill_below, const Phases& ph, const unsigned char* s_nodes, uint32_t stk_base,
                          float4* stage, uint64_t* bars, unsigned char* queue) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    float4* q_o = reinterpret_cast<float4*>(queue);  // origin, max_dist
    float4* q_d = q_o + kQueueCap;                    // direction, pool slot
    const uint32_t stk_ttop = stk_base + (S - 1) * E;
    uint32_t spill[kSpillEntries];
    const int slot_limit = min(pool.capacity, ctl->active_cap);
    Feeder fd;
    feeder_init(fd);
    ShadowRay st;
    st.np = stk_base;
    st.tp = stk_ttop;
    st.nspill = 0;
    st.occluded = false;
    WideCounts wc = {0, 0};
    unsigned wc_tree = 0;
    unsigned dbg_rounds = 0, dbg_has = 0, dbg_nsteps = 0, dbg_tsteps = 0, dbg_chunks = 0;
    bool has = false;
    int slot = -1;
    int qn = 0;
    unsigned rays = 0;
    for (;;) {
        unsigned act = __ballot_sync(0xffffffffu, has);
        while (qn <= kQueueLow && !feeder_exhausted(fd)) {
            feeder_advance<true>(fd, stage, bars, pool.od, pool.sh_d, &ctl->cursor_shadow, slot_limit, lane,
                                 act == 0 && qn == 0);
            if (!fd.fresh) break;
            const float4* buf = stage + fd.cur_buf * (2 * kChunk);
            const float4 o4 = buf[2 * lane], d4 = fd.cur_sh;
            const bool live = o4.w > 0.f;  // a shadow ray waits in this slot
            const int my_slot = fd.cur_base + (int)lane;
            const int verdict = top_shadow(top, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), o4.w, live);
            if (live) rays++;
            // the next-event contribution waits in pend; an occluded ray cancels it
            if (live && verdict == 1) cancel_pend(pool, my_slot);
            const bool tree = live && verdict == 2;
            const unsigned m = __ballot_sync(0xffffffffu, tree);
            if (tree) {
                const int qi = qn + __popc(m & lt_mask);
                q_o[qi] = o4;
                q_d[qi] = make_float4(d4.x, d4.y, d4.z, i2f(my_slot));
            }
            qn += __popc(m);
            if (COUNT && lane == 0) { wc_tree += __popc(m); dbg_chunks++; }
            fd.fresh = false;
            __syncwarp();
        }
        if (qn > 0 && __popc(act) < refill_below) {
            const unsigned idle = ~act;
            const int rank = __popc(idle & lt_mask);
            if (!has && rank < qn) {
                const int qi = qn - 1 - rank;
                const float4 o4 = q_o[qi], d4 = q_d[qi];
                shadow_begin(st, o4, d4, stk_base, stk_ttop, E);
                slot = f2i(d4.w);
                has = true;
            }
            qn = max(0, qn - __popc(idle));
            act = __ballot_sync(0xffffffffu, has);
            __syncwarp();
        }
        if (act == 0) {
            if (qn == 0 && feeder_exhausted(fd)) break;
            continue;
        }
        if (COUNT && lane == 0) { dbg_rounds++; dbg_has += __popc(act); }
#pragma unroll 1
        for (int it = 0; it < ph.node_iters; it++) {
            if (COUNT && lane == 0) dbg_nsteps++;
            shadow_node_step<E, S, COUNT, WIDE>(s_nodes, k_smem, sc, st, stk_base, spill, &wc);
            const unsigned m = __ballot_sync(0xffffffffu, shadow_node_work(st, stk_base) && shadow_has_room<E, S>(st));
            if (__popc(m) < ph.node_min) break;
        }
#pragma unroll 1
        for (;;) {
            if (COUNT && lane == 0) dbg_tsteps++;
            if (PAIR) shadow_tri_step2<E, S, COUNT>(sc, st, stk_base, &wc);
            else shadow_tri_step<E, S, COUNT>(sc, st, stk_base, &wc);
            const unsigned m = __ballot_sync(0xffffffffu, !st.occluded && st.tp != stk_ttop);
            if (__popc(m) < ph.tri_min) break;
        }
        if (has && shadow_done<E, S>(st, stk_base)) {
            // the next-event contribution waits in pend; an occluded ray cancels it
            if (st.occluded) cancel_pend(pool, slot);
            st.np = stk_base;  // an occluded ray leaves entries behind
            st.tp = stk_ttop;
            st.nspill = 0;
            st.occluded = false;
            has = false;
        }
    }
    warp_add(&ctl->cnt_shadow, rays);
    if (COUNT) {
        warp_add(&ctl->cnt_nodes, wc.nodes);
        warp_add(&ctl->cnt_tris, wc.tris);
        warp_add(&ctl->cnt_tree_shadow, wc_tree);
        warp_add(&ctl->dbg[8], dbg_rounds);
        warp_add(&ctl->dbg[9], dbg_has);
        warp_add(&ctl->dbg[10], dbg_nsteps);
        warp_add(&ctl->dbg[11], dbg_tsteps);
        warp_add(&ctl->dbg[12], dbg_chunks);
    }
}

// top of the tree -> shared memory (112-byte stride: the eighth float4 of a node is padding)
template <int THREADS>
TRT_DEV void stage_nodes(unsigned char* s_nodes, const SceneDev& sc, int k_smem) {
    for (int i = threadIdx.x; i < k_smem * 8; i += THREADS)
        if ((i & 7) != 7)
            reinterpret_cast<float4*>(s_nodes)[(i >> 3) * (kSmemNodeStride / 16) + (i & 7)] = __ldg(sc.wide_nodes + i);
}

template <int THREADS, bool COUNT, bool WIDE>
__global__ void __launch_bounds__(THREADS, 1)
k_extend_fast(PoolView pool, SceneDev sc, const __grid_constant__ TopPrims top, Control* ctl, int k_smem,
              int refill_below, Phases ph) {
    constexpr int S = FastCfg<THREADS>::SC;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_nodes = smem;
    uint2* s_stack = reinterpret_cast<uint2*>(smem + (size_t)k_smem * kSmemNodeStride);
    float4* s_stage = reinterpret_cast<float4*>(s_stack + S * THREADS);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_stage) + WARPS * kStageBytesPerWarp);
    unsigned char* s_queue = reinterpret_cast<unsigned char*>(s_bars + WARPS * 2);
    float4* s_top = reinterpret_cast<float4*>(s_queue + WARPS * kQueueBytesClosest);  // (v0|id, e1, e2) per root-level primitive
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    stage_nodes<THREADS>(s_nodes, sc, k_smem);
    if (threadIdx.x < kMaxTop) {
        s_top[threadIdx.x * 3] = top.v0[threadIdx.x];
        s_top[threadIdx.x * 3 + 1] = top.e1[threadIdx.x];
        s_top[threadIdx.x * 3 + 2] = top.e2[threadIdx.x];
    }
    if (lane == 0) {
        mbar_init(&s_bars[warp * 2], 1);
        mbar_init(&s_bars[warp * 2 + 1], 1);
    }
    mbar_fence_init();
    __syncthreads();
    extend_phase<THREADS, S, COUNT, WIDE>(pool, sc, top, ctl, k_smem, refill_below, ph, s_nodes, smem_addr(s_stack + threadIdx.x),
                                          s_stage + warp * (kStageBytesPerWarp / 16), s_bars + warp * 2,
                                          s_queue + warp * kQueueBytesClosest, s_top);
}

template <int THREADS, bool COUNT, bool WIDE>
__global__ void __launch_bounds__(THREADS, 1)
k_shadow_fast(PoolView pool, SceneDev sc, const __grid_constant__ TopPrims top, Control* ctl, int k_smem,
              int refill_below, Phases ph) {
    constexpr int S = FastCfg<THREADS>::SS;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_nodes = smem;
    uint32_t* s_stack = reinterpret_cast<uint32_t*>(smem + (size_t)k_smem * kSmemNodeStride);
    float4* s_stage = reinterpret_cast<float4*>(s_stack + S * THREADS);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_stage) + WARPS * kStageBytesPerWarp);
    unsigned char* s_queue = reinterpret_cast<unsigned char*>(s_bars + WARPS * 2);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    stage_nodes<THREADS>(s_nodes, sc, k_smem);
    if (lane == 0) {
        mbar_init(&s_bars[warp * 2], 1);
        mbar_init(&s_bars[warp * 2 + 1], 1);
    }
    mbar_fence_init();
    __syncthreads();
    shadow_phase<THREADS, S, THREADS * 4, COUNT, WIDE, false>(pool, sc, top, ctl, k_smem, refill_below, ph, s_nodes,
                                                              smem_addr(s_stack + threadIdx.x),
                                                              s_stage + warp * (kStageBytesPerWarp / 16), s_bars + warp * 2,
                                                              s_queue + warp * kQueueBytesShadow);
}

// Shadow rays of the previous shade pass, then the closest-hit rays of this iteration, in ONE persistent launch:
// the two ray sets are independent (both only need that shade pass and this iteration's refill), so a warp that
// runs out of shadow chunks goes straight on to closest-hit chunks -- the tail of the any-hit work, during which
// the stand-alone kernel leaves most of an SM idle, is filled with closest-hit work, and an iteration has one
// launch, one start-up (tree top -> shared memory) and one tail instead of two.  No CTA barrier separates the
// phases: staging buffers, mbarriers (a pair per phase) and queue are per warp, and a thread's any-hit stack
// entries sit in the low words of its own closest-hit entries.
template <int THREADS, bool COUNT, bool WIDE, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
k_trace_fast(PoolView pool, SceneDev sc, const __grid_constant__ TopPrims top, Control* ctl, int k_smem,
             int refill_below, Phases ph_closest, Phases ph_shadow) {
    constexpr int S = FastCfg<THREADS>::SC;
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_nodes = smem;
    uint2* s_stack = reinterpret_cast<uint2*>(smem + (size_t)k_smem * kSmemNodeStride);
    float4* s_stage = reinterpret_cast<float4*>(s_stack + S * THREADS);
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_stage) + WARPS * kStageBytesPerWarp);
    unsigned char* s_queue = reinterpret_cast<unsigned char*>(s_bars + WARPS * 4);
    float4* s_top = reinterpret_cast<float4*>(s_queue + WARPS * kQueueBytesClosest);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (ctl->active_cap <= 0) return;  // the job's last paths were finished by k_finish_paths: nothing to visit
    stage_nodes<THREADS>(s_nodes, sc, k_smem);
    if (threadIdx.x < kMaxTop) {
        s_top[threadIdx.x * 3] = top.v0[threadIdx.x];
        s_top[threadIdx.x * 3 + 1] = top.e1[threadIdx.x];
        s_top[threadIdx.x * 3 + 2] = top.e2[threadIdx.x];
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) mbar_init(&s_bars[warp * 4 + i], 1);
    }
    mbar_fence_init();
    __syncthreads();
    const uint32_t stk_base = smem_addr(s_stack + threadIdx.x);
    float4* stage = s_stage + warp * (kStageBytesPerWarp / 16);
    unsigned char* queue = s_queue + warp * kQueueBytesClosest;
    // any-hit entries (4 bytes): lane l uses bytes [4l, 4l+4) of the 256-byte row its warp owns in every entry plane
    // (the closest-hit entries of the warp's lanes fill the same rows, 8 bytes per lane): conflict free, and private
    // to the warp in both phases
    const uint32_t stk_base_sh = smem_addr(s_stack + (threadIdx.x & ~31u)) + lane * 4;
    shadow_phase<THREADS, S, THREADS * 8, COUNT, WIDE, PAIR>(pool, sc, top, ctl, k_smem, refill_below, ph_shadow, s_nodes, stk_base_sh,
                                                             stage, s_bars + warp * 4, queue);
    __syncwarp();
    extend_phase<THREADS, S, COUNT, WIDE>(pool, sc, top, ctl, k_smem, refill_below, ph_closest, s_nodes, stk_base, stage,
                                          s_bars + warp * 4 + 2, queue, s_top);
}

// ---- drain tail: the last paths of a job run to completion in ONE launch ------------------------
// Near the end of a job the wavefront iterations carry a few ten thousand paths each, and each of them still
// costs three launches with their start-up and tail (the last 25 iterations of a 64-spp C2 job: 2.7 ms for
// under 1 % of the work).  Once the live paths have dropped to `finish_below` and no sample is left, this
// kernel takes every remaining path through extend -> shade -> shadow in a loop, one path per thread, with
// the same device functions the wavefront kernels use (same top phase, same wide-BVH steps, same winner
// verification, same shade_vertex, same order of the radiance additions), so the image is the one the
// remaining iterations would have produced.  The decision is made on the device from Control (every thread reads
// the same values); k_finish_commit then empties the pool bound, which turns the iterations still queued behind
// it into no-ops, and the host's next poll sees alive == 0.
constexpr int kFinishBlock = 128;
constexpr int kFinishStack = 16;

TRT_DEV bool finish_wanted(const Control* ctl, int finish_below) {
    return ctl->active_cap > 0 && ctl->alive <= finish_below && ctl->next_sample >= ctl->total_samples;
}

template <bool COUNT, bool WIDE>
__global__ void __launch_bounds__(kFinishBlock) k_finish_paths(PoolView pool, Control* ctl, SceneDev sc,
                                                               const __grid_constant__ TopPrims top, JobParams job,
                                                               int finish_below) {
    __shared__ uint2 s_stack[kFinishStack * kFinishBlock];
    if (!finish_wanted(ctl, finish_below)) return;
    constexpr uint32_t E = kFinishBlock * 8;
    constexpr int S = kFinishStack;
    const uint32_t stk_base = smem_addr(s_stack + threadIdx.x);
    const uint32_t stk_ttop = stk_base + (S - 1) * E;
    const int cap = ctl->active_cap;
    unsigned n_closest = 0, n_shadow = 0, n_replays = 0;
    WideCounts wc = {0, 0};
    // any-hit query for the lanes with `want`; called by all lanes of the warp
    auto occluded_by = [&](const F3 o, const F3 d, float max_dist, bool want) -> bool {
        __syncwarp();
        const int verdict = top_shadow(top, o, d, max_dist, want);
        if (want) n_shadow++;
        bool occ = want && verdict == 1;
        if (want && verdict == 2) {
            uint32_t spill[kSpillEntries];
            ShadowRay st;
            shadow_begin(st, make_float4(o.x, o.y, o.z, max_dist), make_float4(d.x, d.y, d.z, 0.f), stk_base, stk_ttop, E);
            while (!shadow_done<E, S>(st, stk_base)) {
                shadow_node_step<E, S, COUNT, WIDE>(nullptr, 0, sc, st, stk_base, spill, &wc);
                shadow_tri_step<E, S, COUNT>(sc, st, stk_base, &wc);
            }
            occ = st.occluded;
        }
        __syncwarp();
        return occ;
    };
    for (int first = blockIdx.x * kFinishBlock; first < cap; first += gridDim.x * kFinishBlock) {
        const int slot = first + (int)threadIdx.x;  // cap is a multiple of the block size
        const float4 d4 = pool.od[2 * slot + 1];
        const int flags = f2i(d4.w);
        const int state = flags & 0xff;
        bool active = state == SLOT_ACTIVE;
        const bool live = state != SLOT_DEAD;
        PathVertexIO io;
        io.shadow = false;
        io.shadow_ray.o = io.shadow_ray.d = io.shadow_contrib = f3(0.f, 0.f, 0.f);
        io.shadow_max_dist = 0.f;
        io.depth = (flags >> 8) & 0xff;
        io.prev_mode = (flags >> 16) & 0xff;
        int pix = 0;
        io.thr = f3(1.f, 1.f, 1.f);
        io.rad = f3(0.f, 0.f, 0.f);
        io.ray.o = io.ray.d = f3(0.f, 0.f, 0.f);
        io.rng.v0 = io.rng.v1 = io.rng.v2 = io.rng.v3 = io.rng.v4 = io.rng.d = 0;
        F3 pend = f3(0.f, 0.f, 0.f);
        bool sh = false;
        F3 sh_dir = f3(0.f, 0.f, 0.f);
        float sh_len = 0.f;
        if (live) {
            const float4 o4 = pool.od[2 * slot];
            const uint4 ra = pool.rs[2 * slot], rb = pool.rs[2 * slot + 1];
            const bool fresh = io.depth == 0;  // refill leaves thr / pend unwritten and the pixel in od[2s].w
            io.ray.o = f3(o4.x, o4.y, o4.z);
            io.ray.d = f3(d4.x, d4.y, d4.z);
            if (fresh) {
                pix = f2i(o4.w) & ~kFreshPixelBit;
            } else {
                const float4 thr4 = pool.thr[slot], pend4 = pool.pend[slot];
                pix = f2i(thr4.w);
                io.thr = f3(thr4.x, thr4.y, thr4.z);
                io.rad = f3(__uint_as_float(rb.z), __uint_as_float(rb.w), pend4.w);
                pend = f3(pend4.x, pend4.y, pend4.z);
                sh = o4.w > 0.f;
                if (sh) {
                    const float4 sh4 = pool.sh_d[slot];
                    sh_dir = f3(sh4.x, sh4.y, sh4.z);
                    sh_len = o4.w;
                }
            }
            io.rng.v0 = ra.x; io.rng.v1 = ra.y; io.rng.v2 = ra.z; io.rng.v3 = ra.w;
            io.rng.v4 = rb.x; io.rng.d = rb.y;
        }
        // the shadow ray the last shade pass left behind, then what the next shade pass would do first
        if (occluded_by(io.ray.o, sh_dir, sh_len, sh)) pend = f3(0.f, 0.f, 0.f);
        if (live && io.depth > 0) io.rad = v_add(io.rad, pend);
        bool ended = live && !active;  // SLOT_FINISH
        while (__any_sync(0xffffffffu, active)) {
            float t_hit = 0.f;
            int id = -1;
            if (active) {  // closest hit: top phase, tree phase, winner verification
                n_closest++;
                const TopResult tr = top_closest(top, io.ray.o, io.ray.d);
                t_hit = tr.d_min;
                id = tr.id;
                if (tr.enters) {
                    uint2 spill[kSpillEntries];
                    ClosestRay st;
                    closest_begin(st, make_float4(io.ray.o.x, io.ray.o.y, io.ray.o.z, 0.f),
                                  make_float4(io.ray.d.x, io.ray.d.y, io.ray.d.z, 0.f), tr.d_min, tr.id, stk_base, stk_ttop);
                    while (!closest_done<E, S>(st, stk_base)) {
                        closest_node_step<E, S, COUNT, WIDE>(nullptr, 0, sc, st, stk_base, spill, &wc);
                        closest_tri_step2<E, S, COUNT>(sc, st, stk_base, &wc);
                    }
                    t_hit = st.d_min;
                    id = st.id;
                }
                if (resolve_hit(sc, io.ray.o, io.ray.d, t_hit, id)) n_replays++;
            }
            bool want_shadow = false;
            if (active) {
                if (id < 0 || !shade_vertex(sc, job.rc, io, id, t_hit)) {
                    active = false;  // miss, light source, Russian roulette ...
                    ended = true;
                } else {
                    io.depth++;
                    want_shadow = io.shadow;
                }
            }
            // next-event estimate of this vertex (the wavefront folds it in at the start of the next shade pass)
            const bool occ = occluded_by(io.ray.o, io.shadow_ray.d, io.shadow_max_dist, want_shadow);
            if (want_shadow && !occ) io.rad = v_add(io.rad, io.shadow_contrib);
            if (active && io.depth >= job.rc.max_depth) {  // the reference loop ends after max_depth vertices
                active = false;
                ended = true;
            }
        }
        if (ended) {
            F3 rad = io.rad;
            if (filter_sample(rad)) {  // reference :739-759
                float* a = job.accum + (size_t)pix * 4;
                atomicAdd(a + 0, rad.x);
                atomicAdd(a + 1, rad.y);
                atomicAdd(a + 2, rad.z);
            }
        }
    }
    __syncwarp();
    warp_add(&ctl->cnt_closest, n_closest);
    warp_add(&ctl->cnt_shadow, n_shadow);
    warp_add(&ctl->cnt_replays, n_replays);
    if (COUNT) {
        warp_add(&ctl->cnt_nodes, wc.nodes);
        warp_add(&ctl->cnt_tris, wc.tris);
    }
}

__global__ void k_finish_commit(Control* ctl, int finish_below) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || !finish_wanted(ctl, finish_below)) return;
    ctl->alive = 0;
    ctl->active_cap = 0;
    ctl->compact_go = 0;
}

// ---- 128-byte nodes -> compressed 64-byte nodes (traverse_fast.cuh CNode) -------------------------
// One thread per node.  Per axis: the grid spans the union of the children's boxes with a power-of-two step,
// base is pushed down until plane(0) <= the smallest lo plane, and every child's lo / hi byte is moved outward
// until the plane the TRAVERSAL will decode from it (the same p_fma expression, cnode_plane) contains the
// original plane; if 255 steps do not reach the largest hi plane the step doubles.  The step is also kept
// above four ulps of the largest coordinate so that rounding cannot collapse the grid.
TRT_DEV float cnode_decode(int q, float scale, float base) {
    return p_fma(__uint_as_float(0x4B000000u | (uint32_t)q), scale, base);
}

__global__ void __launch_bounds__(kBlock) k_compress_nodes(const float4* __restrict__ wide, int n, CNode* __restrict__ out,
                                                           int* __restrict__ n_bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* w = wide + (size_t)i * 8;
    const int4 ch = *reinterpret_cast<const int4*>(w + 6);
    const int child[4] = {ch.x, ch.y, ch.z, ch.w};
    CNode c;
    uint32_t qw[6] = {0, 0, 0, 0, 0, 0};
    bool bad = false;
    for (int a = 0; a < 3; a++) {
        const float4 lo4 = w[2 * a], hi4 = w[2 * a + 1];
        const float lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hi[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
        float mn = 3.0e38f, mx = -3.0e38f;
        for (int k = 0; k < 4; k++)
            if (child[k] != kWideEmptyRef) { mn = fminf(mn, lo[k]); mx = fmaxf(mx, hi[k]); }
        if (!(mn <= mx)) { mn = 0.f; mx = 0.f; }  // no children (an empty tree's root)
        // power-of-two step: covers the extent in 254 steps, and at least 4 ulps of the largest coordinate
        const float big = fmaxf(fmaxf(fabsf(mn), fabsf(mx)), 1e-30f);
        int e_min;
        frexpf(big, &e_min);           // big = m * 2^e_min, m in [0.5, 1): ulp(big) = 2^(e_min - 24)
        int e = e_min - 24 + 2;
        const float ext = mx - mn;
        if (ext > 0.f) {
            int ee;
            frexpf(ext / 254.f, &ee);  // ext / 254 <= 2^ee
            e = max(e, ee);
        }
        float scale, base;
        int qlo[4], qhi[4];
        for (int tries = 0;; tries++) {
            scale = ldexpf(1.f, e);
            base = p_fma(-8388608.f, scale, mn);
            for (int it = 0; it < 64 && cnode_decode(0, scale, base) > mn; it++) base = nextafterf(base, -3.0e38f);
            bool ok = cnode_decode(0, scale, base) <= mn && cnode_decode(255, scale, base) >= mx;
            const float origin = cnode_decode(0, scale, base);
            for (int k = 0; k < 4 && ok; k++) {
                if (child[k] == kWideEmptyRef) { qlo[k] = 255; qhi[k] = 0; continue; }
                int ql = min(255, max(0, (int)floorf((lo[k] - origin) / scale)));
                while (ql > 0 && cnode_decode(ql, scale, base) > lo[k]) ql--;
                int qh = min(255, max(0, (int)ceilf((hi[k] - origin) / scale)));
                while (qh < 255 && cnode_decode(qh, scale, base) < hi[k]) qh++;
                ok = cnode_decode(ql, scale, base) <= lo[k] && cnode_decode(qh, scale, base) >= hi[k];
                qlo[k] = ql;
                qhi[k] = qh;
            }
            if (ok) break;
            if (tries == 40) { bad = true; break; }
            e++;
        }
        c.base[a] = base;
        c.scale[a] = scale;
        for (int k = 0; k < 4; k++) {
            qw[2 * a] |= (uint32_t)qlo[k] << (8 * k);
            qw[2 * a + 1] |= (uint32_t)qhi[k] << (8 * k);
        }
    }
    c.qx_lo = qw[0]; c.qx_hi = qw[1]; c.qy_lo = qw[2]; c.qy_hi = qw[3]; c.qz_lo = qw[4]; c.qz_hi = qw[5];
    for (int k = 0; k < 4; k++) c.child[k] = child[k];
    out[i] = c;
    if (bad) atomicAdd(n_bad, 1);
}

// ---- instancing: one parsed mesh, many placements -> the object array the reference's loader would build -------
// out[(i * n_unit + j)] = unit[j] with its three vertices moved to fma(v, scale_i, offset_i): the arithmetic of the
// reference's load_obj (src/loader.cpp:51: v * scale + offset, contracted to one FMA by its -O3 x86 build and
// written as fmaf in host/loader.cpp), so the records equal, byte for byte, what one load_obj call per instance
// produces -- without parsing the file n_instances times and without the array ever existing on the host.
__global__ void __launch_bounds__(kBlock) k_instance_objects(const float4* __restrict__ unit, int n_unit,
                                                             const float4* __restrict__ inst, int n_inst,
                                                             float4* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per float4 of the output
    const long long total = (long long)n_unit * n_inst * 7;
    if (t >= total) return;
    const long long rec = t / 7;
    const int part = (int)(t - rec * 7);
    const int i = (int)(rec / n_unit), j = (int)(rec - (long long)i * n_unit);
    float4 v = __ldg(unit + (size_t)j * 7 + part);
    if (part < 3) {  // v0, v1, v2 (the fourth lane is padding and stays as it is)
        const float4 p = __ldg(inst + i);
        v.x = __fmaf_rn(v.x, p.w, p.x);
        v.y = __fmaf_rn(v.y, p.w, p.y);
        v.z = __fmaf_rn(v.z, p.w, p.z);
    }
    out[t] = v;
}

// ---- parity / test entry points -------------------------------------------------------------
// primary rays of one frame into a scratch pool (slot = reference pixel index)
__global__ void __launch_bounds__(kBlock) k_pack_primary(PoolView pool, JobParams job, float* out_ray) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= pool.capacity) return;
    if (pix >= job.rc.width * job.rc.height) {
        pool.od[2 * pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        pool.od[2 * pix + 1] = make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC)));
        return;
    }
    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
    const int y = job.rc.height - 1 - row;
    Xorwow rng = sample_rng(job, 0, row, col);
    const Ray r = primary_ray(job.cam, col, y, job.rc.width, job.rc.height, rng);
    pool.od[2 * pix] = make_float4(r.o.x, r.o.y, r.o.z, 0.f);
    pool.od[2 * pix + 1] = make_float4(r.d.x, r.d.y, r.d.z, i2f(pack_flags(SLOT_ACTIVE, 0, MODE_SPEC)));
    if (out_ray) {
        float* p = out_ray + (size_t)pix * 6;
        p[0] = r.o.x; p[1] = r.o.y; p[2] = r.o.z; p[3] = r.d.x; p[4] = r.d.y; p[5] = r.d.z;
    }
}

// caller-provided rays (8 floats: o.xyz, d.xyz, t_max, unused) into a scratch pool, as closest-hit
// rays and as shadow rays at once (a shadow ray needs t_max > 0: that word is also the "trace it" flag)
__global__ void __launch_bounds__(kBlock) k_pack_rays(PoolView pool, const float* __restrict__ rays, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.capacity) return;
    if (i >= n) {
        pool.od[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pool.od[2 * i + 1] = make_float4(0.f, 0.f, 0.f, i2f(pack_flags(SLOT_DEAD, 0, MODE_SPEC)));
        return;
    }
    const float* p = rays + (size_t)i * 8;
    pool.od[2 * i] = make_float4(p[0], p[1], p[2], p[6]);
    pool.od[2 * i + 1] = make_float4(p[3], p[4], p[5], i2f(pack_flags(SLOT_ACTIVE, 0, MODE_SPEC)));
    pool.sh_d[i] = make_float4(p[3], p[4], p[5], 0.f);
    pool.pend[i] = make_float4(1.f, 1.f, 1.f, 0.f);
}

// hits of the FAST kernel over a scratch pool -> caller's arrays, through the same winner
// verification / replay the shade kernel applies; `out_replayed` marks the replayed rays
__global__ void __launch_bounds__(kBlock) k_unpack_hits(PoolView pool, SceneDev sc, int n, int* out_id, float* out_t,
                                                        int* out_replayed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 h = pool.hit[i];
    const float4 o4 = pool.od[2 * i], d4 = pool.od[2 * i + 1];
    float t = h.x;
    int id = f2i(h.y);
    const bool replayed = resolve_hit(sc, f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z), t, id);
    if (out_id) out_id[i] = id;
    if (out_t) out_t[i] = t;
    if (out_replayed) out_replayed[i] = replayed ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock) k_unpack_occluded(PoolView pool, int n, int* out_occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_occ[i] = pool.pend[i].x == 0.f ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock) k_trace_primary_ref(SceneDev sc, JobParams job, int* out_id, float* out_t,
                                                              float* out_ray, uint32_t* out_fetched,
                                                              uint32_t* out_entered, uint32_t* out_tris) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= job.rc.width * job.rc.height) return;
    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
    const int y = job.rc.height - 1 - row;
    Xorwow rng = sample_rng(job, 0, row, col);
    const Ray r = primary_ray(job.cam, col, y, job.rc.width, job.rc.height, rng);
    float t;
    VisitCounts vc = {0, 0, 0};
    const int id = ref_closest<true>(sc, r, &t, &vc);
    if (out_id) out_id[pix] = id;
    if (out_t) out_t[pix] = t;
    if (out_ray) {
        float* p = out_ray + (size_t)pix * 6;
        p[0] = r.o.x; p[1] = r.o.y; p[2] = r.o.z; p[3] = r.d.x; p[4] = r.d.y; p[5] = r.d.z;
    }
    if (out_fetched) out_fetched[pix] = vc.fetched;
    if (out_entered) out_entered[pix] = vc.entered;
    if (out_tris) out_tris[pix] = vc.tris;
}

__global__ void __launch_bounds__(kBlock) k_trace_closest_ref(SceneDev sc, const float* __restrict__ rays, int n,
                                                              int* out_id, float* out_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 8;
    Ray r;
    r.o = f3(p[0], p[1], p[2]);
    r.d = f3(p[3], p[4], p[5]);
    float t;
    VisitCounts vc = {0, 0, 0};
    out_id[i] = ref_closest<false>(sc, r, &t, &vc);
    if (out_t) out_t[i] = t;
}

__global__ void __launch_bounds__(kBlock) k_trace_shadow_ref(SceneDev sc, const float* __restrict__ rays, int n,
                                                             int* out_occ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 8;
    Ray r;
    r.o = f3(p[0], p[1], p[2]);
    r.d = f3(p[3], p[4], p[5]);
    VisitCounts vc = {0, 0, 0};
    out_occ[i] = ref_shadow<false>(sc, r, p[6], &vc) ? 1 : 0;
}

__global__ void k_rng_states(JobParams job, int f, int first_pixel, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pix = first_pixel + i;
    const int row = pix / job.rc.width, col = pix - row * job.rc.width;
    const Xorwow s = sample_rng(job, f, row, col);
    uint32_t* o = out + (size_t)i * 6;
    o[0] = s.v0; o[1] = s.v1; o[2] = s.v2; o[3] = s.v3; o[4] = s.v4; o[5] = s.d;
}

// accum/frames -> gamma 2.2 -> ARGB8888 (reference src/pipeline.cpp:59-71, common.h:114-128)
__global__ void k_tonemap(const float4* __restrict__ accum, int n, float inv_frames, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = accum[i];
    const float c[3] = {a.x * inv_frames, a.y * inv_frames, a.z * inv_frames};
    int q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float v = c[k] < 0.f ? 0.f : (c[k] > 1.f ? 1.f : c[k]);
        q[k] = (int)(pow((double)v, 1 / 2.2) * 255 + .5);
    }
    out[i] = (255u << 24) | ((uint32_t)q[0] << 16) | ((uint32_t)q[1] << 8) | (uint32_t)q[2];
}

int grid_for(int n) { return (n + kBlock - 1) / kBlock; }

template <int THREADS>
size_t fast_smem_bytes(int k_smem, bool shadow) {
    const size_t stack = shadow ? (size_t)FastCfg<THREADS>::SS * THREADS * 4 : (size_t)FastCfg<THREADS>::SC * THREADS * 8;
    return (size_t)k_smem * kSmemNodeStride + stack + (shadow ? 0 : kTopStageBytes) +
           (size_t)(THREADS / 32) * (kStageBytesPerWarp + 16 + (shadow ? kQueueBytesShadow : kQueueBytesClosest));
}

template <int THREADS>
size_t trace_smem_bytes(int k_smem) {  // the combined kernel: closest-hit layout with two more mbarriers per warp
    return fast_smem_bytes<THREADS>(k_smem, false) + (size_t)(THREADS / 32) * 16;
}

template <int THREADS, bool COUNT>
void launch_trace_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl,
                       const LaunchDims& dims, cudaStream_t s) {
    if (dims.wide_loads && THREADS == 896) {
        k_trace_fast<896, COUNT, true, false><<<dims.sms, 896, trace_smem_bytes<896>(0), s>>>(
            pool, sc, top, ctl, 0, dims.refill_below, dims.closest_phases, dims.shadow_phases);
        return;
    }
    // the two extra mbarriers per warp come out of the staged nodes
    const int extra_nodes = (int)(((size_t)(THREADS / 32) * 16 + kSmemNodeStride - 1) / kSmemNodeStride);
    const int k = max(0, min(dims.smem_nodes - extra_nodes, sc.n_wide_nodes));
    k_trace_fast<THREADS, COUNT, false, false><<<dims.sms, THREADS, trace_smem_bytes<THREADS>(k), s>>>(
        pool, sc, top, ctl, k, dims.refill_below, dims.closest_phases, dims.shadow_phases);
}
template <bool COUNT>
void trace_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl, const LaunchDims& dims,
                cudaStream_t s) {
    switch (dims.fast_threads) {
    case 1024: launch_trace_fast<1024, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 896: launch_trace_fast<896, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 768: launch_trace_fast<768, COUNT>(pool, sc, top, ctl, dims, s); break;
    default: launch_trace_fast<512, COUNT>(pool, sc, top, ctl, dims, s); break;
    }
}

template <int THREADS, bool COUNT>
void launch_extend_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl,
                        const LaunchDims& dims, cudaStream_t s) {
    if (dims.wide_loads && THREADS == 896) {  // big scenes: compressed nodes, nothing staged
        k_extend_fast<896, COUNT, true><<<dims.sms, 896, fast_smem_bytes<896>(0, false), s>>>(
            pool, sc, top, ctl, 0, dims.refill_below, dims.closest_phases);
        return;
    }
    const int k = min(dims.smem_nodes, sc.n_wide_nodes);
    k_extend_fast<THREADS, COUNT, false><<<dims.sms, THREADS, fast_smem_bytes<THREADS>(k, false), s>>>(
        pool, sc, top, ctl, k, dims.refill_below, dims.closest_phases);
}
template <int THREADS, bool COUNT>
void launch_shadow_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl,
                        const LaunchDims& dims, cudaStream_t s) {
    if (dims.wide_loads && THREADS == 896) {
        k_shadow_fast<896, COUNT, true><<<dims.sms, 896, fast_smem_bytes<896>(0, true), s>>>(
            pool, sc, top, ctl, 0, dims.refill_below, dims.shadow_phases);
        return;
    }
    const int k = min(dims.smem_nodes, sc.n_wide_nodes);
    k_shadow_fast<THREADS, COUNT, false><<<dims.sms, THREADS, fast_smem_bytes<THREADS>(k, true), s>>>(
        pool, sc, top, ctl, k, dims.refill_below, dims.shadow_phases);
}

template <bool COUNT>
void extend_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl, const LaunchDims& dims,
                 cudaStream_t s) {
    switch (dims.fast_threads) {
    case 1024: launch_extend_fast<1024, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 896: launch_extend_fast<896, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 768: launch_extend_fast<768, COUNT>(pool, sc, top, ctl, dims, s); break;
    default: launch_extend_fast<512, COUNT>(pool, sc, top, ctl, dims, s); break;
    }
}
template <bool COUNT>
void shadow_fast(const PoolView& pool, const SceneDev& sc, const TopPrims& top, Control* ctl, const LaunchDims& dims,
                 cudaStream_t s) {
    switch (dims.fast_threads) {
    case 1024: launch_shadow_fast<1024, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 896: launch_shadow_fast<896, COUNT>(pool, sc, top, ctl, dims, s); break;
    case 768: launch_shadow_fast<768, COUNT>(pool, sc, top, ctl, dims, s); break;
    default: launch_shadow_fast<512, COUNT>(pool, sc, top, ctl, dims, s); break;
    }
}

template <class K>
int opt_in_smem(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess ? 0 : -1;
}

}  // namespace

// ---- launchers -------------------------------------------------------------------------------
int wf_configure() {
    int rc = 0;

    rc |= opt_in_smem(k_extend_fast<512, false, false>);
    rc |= opt_in_smem(k_extend_fast<512, true, false>);
    rc |= opt_in_smem(k_extend_fast<768, false, false>);
    rc |= opt_in_smem(k_extend_fast<768, true, false>);
    rc |= opt_in_smem(k_extend_fast<896, false, false>);
    rc |= opt_in_smem(k_extend_fast<896, true, false>);
    rc |= opt_in_smem(k_shadow_fast<896, false, false>);
    rc |= opt_in_smem(k_shadow_fast<896, true, false>);
    rc |= opt_in_smem(k_trace_fast<896, false, false, false>);
    rc |= opt_in_smem(k_trace_fast<896, true, false, false>);
    rc |= opt_in_smem(k_extend_fast<1024, false, false>);
    rc |= opt_in_smem(k_extend_fast<1024, true, false>);
    rc |= opt_in_smem(k_shadow_fast<512, false, false>);
    rc |= opt_in_smem(k_shadow_fast<512, true, false>);
    rc |= opt_in_smem(k_shadow_fast<768, false, false>);
    rc |= opt_in_smem(k_shadow_fast<768, true, false>);
    rc |= opt_in_smem(k_shadow_fast<1024, false, false>);
    rc |= opt_in_smem(k_shadow_fast<1024, true, false>);
    rc |= opt_in_smem(k_extend_fast<896, false, true>);
    rc |= opt_in_smem(k_extend_fast<896, true, true>);
    rc |= opt_in_smem(k_shadow_fast<896, false, true>);
    rc |= opt_in_smem(k_shadow_fast<896, true, true>);
    rc |= opt_in_smem(k_trace_fast<512, false, false, false>);
    rc |= opt_in_smem(k_trace_fast<512, true, false, false>);
    rc |= opt_in_smem(k_trace_fast<768, false, false, false>);
    rc |= opt_in_smem(k_trace_fast<768, true, false, false>);
    rc |= opt_in_smem(k_trace_fast<1024, false, false, false>);
    rc |= opt_in_smem(k_trace_fast<1024, true, false, false>);
    rc |= opt_in_smem(k_trace_fast<896, false, true, false>);
    rc |= opt_in_smem(k_trace_fast<896, true, true, false>);
    return rc;
}

size_t wf_fast_smem_bytes(int threads, int smem_nodes, bool shadow) {
    switch (threads) {
    case 512: return fast_smem_bytes<512>(smem_nodes, shadow);
    case 768: return fast_smem_bytes<768>(smem_nodes, shadow);
    case 896: return fast_smem_bytes<896>(smem_nodes, shadow);
    case 1024: return fast_smem_bytes<1024>(smem_nodes, shadow);
    default: return 0;
    }
}

int wf_fast_max_smem_nodes(int threads, size_t smem_limit) {
    const size_t fixed = wf_fast_smem_bytes(threads, 0, false);  // the closest-hit kernel has the larger stacks
    if (fixed == 0 || fixed >= smem_limit) return 0;
    return (int)((smem_limit - fixed) / kSmemNodeStride);
}

void wf_init_pool(const PoolView& pool, cudaStream_t s) {
    k_init_pool<<<grid_for(pool.capacity), kBlock, 0, s>>>(pool);
}

void wf_reset_counters(Control* ctl, cudaStream_t s) { k_reset_counters<<<1, 32, 0, s>>>(ctl); }

void wf_begin_job(Control* ctl, unsigned long long total_samples, int pool_capacity, cudaStream_t s) {
    k_begin_job<<<1, 32, 0, s>>>(ctl, total_samples, pool_capacity);
}

void wf_col_table(const uint4* col_a, const uint32_t* col_b, int w, int first_frame_seed, int frame_stride,
                  int seed_base, int n_frames, XwColVec* out, cudaStream_t s) {
    const long long n = (long long)n_frames * w;
    k_col_table<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(col_a, col_b, w, first_frame_seed, frame_stride, seed_base, n_frames,
                                                            out);
}

// One iteration on one stream: refill -> [compact] -> trace -> shade.
//   refill : k_refill (free scan + bookkeeping + regeneration of the slots that ended in the last shade pass)
//   compact: drain phase only (the host launches the three kernels once the job is close to its drain phase;
//            k_refill lets a compaction go only in an iteration that regenerates nothing)
//   trace  : FAST, combined -- the shadow rays the last shade pass wrote, then this iteration's closest-hit rays,
//            one persistent launch (k_trace_fast).  FAST with merged_trace off, and REF: closest hit before
//            shade, any hit after it, two launches
//   shade  : one thread per slot
//   finish : drain tail (combined mode): k_finish_paths + commit ride along once few paths are left
template <int MODE, bool COUNT>
static int iteration_impl(const PoolView& pool, Control* ctl, const SceneDev& sc, const TopPrims& top,
                          const JobParams& job, const LaunchDims& dims, const IterStreams& st, cudaEvent_t* marks,
                          int* compact_lists) {
    const int full = pool.capacity / kBlock;
    const int persistent = dims.sms * 8;
    cudaStream_t s = st.main;
    int launched = 0;
    // marks[0..5]; st.mark_mask selects which of them are recorded
    auto mark = [&](int i) { if (marks && ((st.mark_mask >> i) & 1)) cudaEventRecord(marks[i], s); };
    mark(0);
    const int visit = min(pool.capacity, st.visit_cap);
    k_refill<<<((visit + 31) / 32 + kRefillBlock - 1) / kRefillBlock, kRefillBlock, 0, s>>>(
        pool, ctl, job, dims.compact_quarters, st.samples_left ? 1 : 0, dims.debug_checks ? 1 : 0);
    launched += 1;
    mark(1);
    if (compact_lists && visit > kCompactMinCap) {
        int* list_a = compact_lists;
        int* list_b = compact_lists + pool.capacity + 512;  // each list is sized for the whole pool
        k_compact_scan<<<dims.sms * 2, 1024, 0, s>>>(pool, ctl, list_a, list_b);
        k_compact_move<<<persistent, kBlock, 0, s>>>(pool, ctl, list_a, list_b, dims.debug_checks ? 1 : 0);
        k_compact_commit<<<1, 32, 0, s>>>(ctl);
        launched += 3;
    }
    mark(2);
    const bool merged = MODE == TRT_TRAVERSE_FAST && dims.merged_trace;
    if (merged) trace_fast<COUNT>(pool, sc, top, ctl, dims, s);
    else if (MODE == TRT_TRAVERSE_FAST) extend_fast<COUNT>(pool, sc, top, ctl, dims, s);
    else k_extend_ref<COUNT><<<full, kBlock, 0, s>>>(pool, sc, ctl);
    mark(3);
    // blocks beyond active_cap return at once, but 16 Ki of them still cost 0.1 ms: size the grid by the bound
    const int shade_blocks = (visit + dims.shade_block - 1) / dims.shade_block;
    constexpr bool F = MODE == TRT_TRAVERSE_FAST;
    const int eager = (st.samples_left || st.mostly_live) ? 1 : 0;
    if (dims.shade_block == 128) {
        switch (dims.shade_minb) {
        case 9: k_shade<COUNT, F, 128, 9><<<shade_blocks, 128, 0, s>>>(pool, ctl, sc, job, eager); break;
        case 10: k_shade<COUNT, F, 128, 10><<<shade_blocks, 128, 0, s>>>(pool, ctl, sc, job, eager); break;
        case 12: k_shade<COUNT, F, 128, 12><<<shade_blocks, 128, 0, s>>>(pool, ctl, sc, job, eager); break;
        case 14: k_shade<COUNT, F, 128, 14><<<shade_blocks, 128, 0, s>>>(pool, ctl, sc, job, eager); break;
        default: k_shade<COUNT, F, 128, 8><<<shade_blocks, 128, 0, s>>>(pool, ctl, sc, job, eager); break;
        }
    } else {
        k_shade<COUNT, F, kShadeMaxBlock, 2><<<shade_blocks, dims.shade_block, 0, s>>>(pool, ctl, sc, job, eager);
    }
    launched += 2;
    mark(4);
    if (!merged) {
        if (MODE == TRT_TRAVERSE_FAST) shadow_fast<COUNT>(pool, sc, top, ctl, dims, s);
        else k_shadow_ref<COUNT><<<full, kBlock, 0, s>>>(pool, sc, ctl);
        launched += 1;
    }
    mark(5);
    if (merged && st.finish_below > 0) {
        // drain tail: once few enough paths are left they are run to completion in one launch (decided on the device)
        const int blocks = max(1, min((visit + kFinishBlock - 1) / kFinishBlock, dims.sms * 16));
        if (dims.wide_loads) k_finish_paths<COUNT, true><<<blocks, kFinishBlock, 0, s>>>(pool, ctl, sc, top, job, st.finish_below);
        else k_finish_paths<COUNT, false><<<blocks, kFinishBlock, 0, s>>>(pool, ctl, sc, top, job, st.finish_below);
        k_finish_commit<<<1, 32, 0, s>>>(ctl, st.finish_below);
        launched += 2;
    }
    return launched;
}

int wf_iteration(const PoolView& pool, Control* ctl, const SceneDev& sc, const TopPrims& top,
                 const JobParams& job, int traversal, bool count, const LaunchDims& dims, const IterStreams& st,
                 cudaEvent_t* marks, int* compact_lists) {
    if (traversal == TRT_TRAVERSE_REF) {
        if (count) return iteration_impl<TRT_TRAVERSE_REF, true>(pool, ctl, sc, top, job, dims, st, marks, compact_lists);
        return iteration_impl<TRT_TRAVERSE_REF, false>(pool, ctl, sc, top, job, dims, st, marks, compact_lists);
    }
    if (count) return iteration_impl<TRT_TRAVERSE_FAST, true>(pool, ctl, sc, top, job, dims, st, marks, compact_lists);
    return iteration_impl<TRT_TRAVERSE_FAST, false>(pool, ctl, sc, top, job, dims, st, marks, compact_lists);
}

void wf_trace_primary(const SceneDev& sc, const JobParams& job, int, int traversal, int* d_id, float* d_t,
                      float* d_ray, uint32_t* d_fetched, uint32_t* d_entered, uint32_t* d_tris,
                      const TopPrims& top, const PoolView& scratch, Control* ctl, const LaunchDims& dims,
                      cudaStream_t s) {
    const int n = job.rc.width * job.rc.height;
    if (traversal == TRT_TRAVERSE_REF) {
        k_trace_primary_ref<<<grid_for(n), kBlock, 0, s>>>(sc, job, d_id, d_t, d_ray, d_fetched, d_entered, d_tris);
        return;
    }
    // FAST: the production persistent kernel over a scratch pool; "entered" reports replayed rays
    k_pack_primary<<<grid_for(scratch.capacity), kBlock, 0, s>>>(scratch, job, d_ray);
    k_reset_cursors<<<1, 32, 0, s>>>(ctl);
    extend_fast<false>(scratch, sc, top, ctl, dims, s);
    k_unpack_hits<<<grid_for(n), kBlock, 0, s>>>(scratch, sc, n, d_id, d_t, reinterpret_cast<int*>(d_entered));
    (void)d_fetched;
    (void)d_tris;
}

void wf_trace_closest(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_id, float* d_t,
                      const TopPrims& top, const PoolView& scratch, Control* ctl, const LaunchDims& dims,
                      cudaStream_t s) {
    if (traversal == TRT_TRAVERSE_REF) {
        k_trace_closest_ref<<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_id, d_t);
        return;
    }
    k_pack_rays<<<grid_for(scratch.capacity), kBlock, 0, s>>>(scratch, d_rays, n);
    k_reset_cursors<<<1, 32, 0, s>>>(ctl);
    extend_fast<false>(scratch, sc, top, ctl, dims, s);
    k_unpack_hits<<<grid_for(n), kBlock, 0, s>>>(scratch, sc, n, d_id, d_t, nullptr);
}

void wf_trace_shadow(const SceneDev& sc, const float* d_rays, int n, int traversal, int* d_occ,
                     const TopPrims& top, const PoolView& scratch, Control* ctl, const LaunchDims& dims,
                     cudaStream_t s) {
    if (traversal == TRT_TRAVERSE_REF) {
        k_trace_shadow_ref<<<grid_for(n), kBlock, 0, s>>>(sc, d_rays, n, d_occ);
        return;
    }
    k_pack_rays<<<grid_for(scratch.capacity), kBlock, 0, s>>>(scratch, d_rays, n);
    k_reset_cursors<<<1, 32, 0, s>>>(ctl);
    shadow_fast<false>(scratch, sc, top, ctl, dims, s);
    k_unpack_occluded<<<grid_for(n), kBlock, 0, s>>>(scratch, n, d_occ);
}

void wf_rng_states(const JobParams& job, int frame_local, int first_pixel, int n, uint32_t* d_states,
                   cudaStream_t s) {
    k_rng_states<<<grid_for(n), kBlock, 0, s>>>(job, frame_local, first_pixel, n, d_states);
}

// 128-byte wide nodes -> compressed 64-byte nodes; *d_bad counts nodes whose boxes could not be contained (never
// seen; the caller then keeps the uncompressed nodes)
void wf_compress_nodes(const float4* d_wide, int n, uint4* d_cnodes, int* d_bad, cudaStream_t s) {
    if (n > 0) k_compress_nodes<<<grid_for(n), kBlock, 0, s>>>(d_wide, n, reinterpret_cast<CNode*>(d_cnodes), d_bad);
}

void wf_instance_objects(const float4* d_unit, int n_unit, const float4* d_inst, int n_inst, float4* d_out, cudaStream_t s) {
    const long long total = (long long)n_unit * n_inst * 7;
    k_instance_objects<<<(unsigned)((total + kBlock - 1) / kBlock), kBlock, 0, s>>>(d_unit, n_unit, d_inst, n_inst, d_out);
}

void wf_tonemap(const float* d_accum, int n_pixels, int frames, uint32_t* d_argb, cudaStream_t s) {
    k_tonemap<<<grid_for(n_pixels), kBlock, 0, s>>>(reinterpret_cast<const float4*>(d_accum), n_pixels,
                                                    1.0f / frames, d_argb);
}

}  // namespace trt
