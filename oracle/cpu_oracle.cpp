// cpu_oracle.cpp -- TEST INFRASTRUCTURE (oracle).  Builds into oracle/_build/liboracle_cpu.so.
//
// CPU restatement of the reference's path tracer: the per-sample algorithm of
// oracle/pt_restatement.h (reference src/renderer.cu:317-760) driven per pixel with the
// XORWOW streams of oracle/xorwow_ref.h, OpenMP over pixels.  The reference itself has no
// CPU renderer; this is the "CPU path" baseline of BASELINE.md section 3 (labelled a
// restatement) and a GPU-less sanity oracle.  It is NOT bit-identical to a GPU run
// (MUFU approximations, FTZ, FMA contraction) -- the parity oracle proper is the unmodified
// reference kernel in oracle/_ref.  Never linked into or loaded by the product.
#include "pt_restatement.h"
#include "xorwow_ref.h"

#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace {

struct HostRng {
    xwref::State st;
    float uniform() { return xwref::uniform(&st); }
};

struct HostTexture {
    int w, h;
    const unsigned char* rgb;
};

// tex2D<float4> on an RGBA8 array: wrap addressing, normalised coordinates, bilinear filter
// with 8-bit fractional weights (CUDA programming guide, "Linear Filtering").
struct HostTex {
    const HostTexture* t;
    int n;
    static float frac8(float x) { return floorf(x * 256.f + 0.5f) / 256.f; }
    ptr::V3 operator()(int id, float u, float v) const {
        if (id < 0 || id >= n || !t[id].rgb) return ptr::mk(0, 0, 0);
        const HostTexture& T = t[id];
        u -= floorf(u);
        v -= floorf(v);
        float x = u * T.w - 0.5f, y = v * T.h - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float a = frac8(x - fx), b = frac8(y - fy);
        int x0 = ((int)fx % T.w + T.w) % T.w, y0 = ((int)fy % T.h + T.h) % T.h;
        int x1 = (x0 + 1) % T.w, y1 = (y0 + 1) % T.h;
        float c[3];
        for (int k = 0; k < 3; k++) {
            float t00 = T.rgb[(y0 * T.w + x0) * 3 + k] / 255.f, t10 = T.rgb[(y0 * T.w + x1) * 3 + k] / 255.f;
            float t01 = T.rgb[(y1 * T.w + x0) * 3 + k] / 255.f, t11 = T.rgb[(y1 * T.w + x1) * 3 + k] / 255.f;
            c[k] = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
        }
        return ptr::mk(c[0], c[1], c[2]);
    }
};

ptr::SceneView view(const void* objects, const void* nodes, const int* lights, int n_lights) {
    ptr::SceneView sc;
    sc.objects = (const ptr::ObjRec*)objects;
    sc.nodes = (const ptr::NodeRec*)nodes;
    sc.lights = lights;
    sc.light_count = n_lights;
    return sc;
}

}  // namespace

extern "C" {

int oracle_xorwow_state(unsigned long long seed, unsigned long long subsequence, unsigned* out) {
    xwref::State s;
    xwref::init(seed, subsequence, &s);
    for (int k = 0; k < 5; k++) out[k] = s.v[k];
    out[5] = s.d;
    return 0;
}

int oracle_xorwow_draws(unsigned long long seed, unsigned long long subsequence, int n, unsigned* u32, float* f32) {
    xwref::State a, b;
    xwref::init(seed, subsequence, &a);
    b = a;
    for (int i = 0; i < n; i++) {
        if (u32) u32[i] = xwref::next(&a);
        if (f32) f32[i] = xwref::uniform(&b);
    }
    return 0;
}

// Primary rays of one frame (reference renderer.cu:319-425).  Outputs are host arrays indexed by
// the reference pixel index i = (h-1-y)*w + x; any may be NULL.  Rows [row0, row1) of the image
// (in i / w terms) are computed, so callers can bound the work.
int oracle_primary(const void* objects, const void* nodes, const void* cam, int w, int h, int seed, int row0,
                   int row1, int* ids, float* ts, float* rays, unsigned* fetched, unsigned* entered, unsigned* tris,
                   int threads) {
    ptr::SceneView sc = view(objects, nodes, nullptr, 0);
    ptr::CamRec c;
    memcpy(&c, cam, sizeof(c));
    xwref::sequence_matrices();
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = (long long)row0 * w; i < (long long)row1 * w; i++) {
        int x = (int)(i % w), y = h - 1 - (int)(i / w);
        HostRng rng;
        xwref::init((unsigned long long)seed, (unsigned long long)i, &rng.st);
        ptr::V3 ro, rd;
        ptr::primary_ray(c, x, y, w, h, rng, &ro, &rd);
        ptr::Counters cnt = {0, 0, 0, 0, 0};
        float t;
        int id = ptr::closest_hit(sc, ro, rd, &t, &cnt);
        if (ids) ids[i] = id;
        if (ts) ts[i] = t;
        if (rays) { float* r = rays + 6 * i; r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z; }
        if (fetched) fetched[i] = (unsigned)cnt.nodes_fetched;
        if (entered) entered[i] = (unsigned)cnt.nodes_entered;
        if (tris) tris[i] = (unsigned)cnt.tris_tested;
    }
    return 0;
}

// Full render: adds frames first_frame .. first_frame+n_frames-1 (seed = seed_base + frame) into
// accum (w*h records of 4 floats) for image rows [row0,row1).  tex: n_tex RGB8 images given as
// parallel arrays.  totals[5] (may be NULL): closest rays, shadow rays, nodes fetched, entered, tris.
int oracle_render(const void* objects, const void* nodes, const int* lights, int n_lights, const void* cam, int w,
                  int h, int seed_base, int first_frame, int n_frames, int max_depth, int rr_threshold,
                  const unsigned char* const* tex_rgb, const int* tex_w, const int* tex_h, int n_tex, int row0,
                  int row1, float* accum, unsigned long long* totals, int threads) {
    ptr::SceneView sc = view(objects, nodes, lights, n_lights);
    ptr::CamRec c;
    memcpy(&c, cam, sizeof(c));
    ptr::Consts k = {max_depth, rr_threshold};
    std::vector<HostTexture> tv(n_tex > 0 ? n_tex : 1);
    for (int i = 0; i < n_tex; i++) tv[i] = HostTexture{tex_w[i], tex_h[i], tex_rgb[i]};
    HostTex tex{tv.data(), n_tex};
    xwref::sequence_matrices();
    if (threads > 0) omp_set_num_threads(threads);
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : t0, t1, t2, t3, t4)
    for (long long i = (long long)row0 * w; i < (long long)row1 * w; i++) {
        int x = (int)(i % w), y = h - 1 - (int)(i / w);
        for (int f = 0; f < n_frames; f++) {
            HostRng rng;
            xwref::init((unsigned long long)(seed_base + first_frame + f), (unsigned long long)i, &rng.st);
            ptr::V3 ro, rd, rad;
            ptr::primary_ray(c, x, y, w, h, rng, &ro, &rd);
            ptr::Counters cnt = {0, 0, 0, 0, 0};
            if (ptr::trace_sample(sc, k, ro, rd, rng, tex, &rad, &cnt)) {
                accum[4 * i + 0] += rad.x;
                accum[4 * i + 1] += rad.y;
                accum[4 * i + 2] += rad.z;
            }
            t0 += cnt.closest_rays; t1 += cnt.shadow_rays; t2 += cnt.nodes_fetched; t3 += cnt.nodes_entered;
            t4 += cnt.tris_tested;
        }
    }
    if (totals) { totals[0] = t0; totals[1] = t1; totals[2] = t2; totals[3] = t3; totals[4] = t4; }
    return 0;
}

int oracle_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
