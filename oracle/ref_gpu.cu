// ref_gpu.cu -- TEST INFRASTRUCTURE (oracle).  Builds into oracle/_ref/libtrt_ref.so.
//
// This translation unit textually includes the UNMODIFIED reference renderer
// (/root/reference/src/renderer.cu, given as REF_RENDERER_CU on the command line)
// and is compiled with the reference's own nvcc flags (Makefile:48,55, -arch raised
// to sm_100).  Around it sit
//   * a C ABI so tests/bench can drive the reference's own entry points
//     (init_scene_data, launch_render_kernel, load_obj, BVH::build, create_cornell_box,
//     CameraController::get_params) from Python through ctypes;
//   * the "ID as emission" first-hit oracle: the unmodified kernel renders a
//     copy of the scene whose triangle k emits a colour encoding k, so one launch on a
//     zeroed buffer yields the reference's first-hit id per pixel;
//   * instrumented kernels built from oracle/pt_restatement.h that report what the
//     unmodified kernel cannot: primary rays, d_min, node-visit counters, ray counts.
// Nothing here is used by the product library.
#include REF_RENDERER_CU

#include "loader.h"
#include "camera.h"
#include "image_io.h"
#include "pipeline.h"
#include "pt_restatement.h"

#include <chrono>
#include <cstdlib>
#include <vector>
#include <string>

static_assert(sizeof(ptr::ObjRec) == sizeof(Object), "restatement layout");
static_assert(sizeof(ptr::NodeRec) == sizeof(LinearBVHNode), "restatement layout");
static_assert(sizeof(ptr::CamRec) == sizeof(CameraParams), "restatement layout");

namespace {

struct CurandRng {
    curandState st;
    __device__ float uniform() { return curand_uniform(&st); }
};

struct DeviceTex {
    __device__ ptr::V3 operator()(int id, float u, float v) const {
        float4 t = tex2D<float4>(d_textures[id], u, v);
        return ptr::mk(t.x, t.y, t.z);
    }
};

__global__ void k_primary_counts(int width, int height, int seed, ptr::CamRec cam, ptr::SceneView sc,
                                 int* out_id, float* out_t, float* out_ray, unsigned* out_fetched,
                                 unsigned* out_entered, unsigned* out_tris) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    int i = (height - y - 1) * width + x;  // renderer.cu:322
    CurandRng rng;
    curand_init(seed, i, 0, &rng.st);      // renderer.cu:326
    ptr::V3 ro, rd;
    ptr::primary_ray(cam, x, y, width, height, rng, &ro, &rd);
    ptr::Counters c = {0, 0, 0, 0, 0};
    float t;
    int id = ptr::closest_hit(sc, ro, rd, &t, &c);
    if (out_id) out_id[i] = id;
    if (out_t) out_t[i] = t;
    if (out_ray) {
        float* r = out_ray + 6 * (size_t)i;
        r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z;
    }
    if (out_fetched) out_fetched[i] = (unsigned)c.nodes_fetched;
    if (out_entered) out_entered[i] = (unsigned)c.nodes_entered;
    if (out_tris) out_tris[i] = (unsigned)c.tris_tested;
}

__global__ void k_full_counts(Vec* accum, int width, int height, int seed, ptr::CamRec cam, ptr::SceneView sc,
                              ptr::Consts k, unsigned long long* totals) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    int i = (height - y - 1) * width + x;
    CurandRng rng;
    curand_init(seed, i, 0, &rng.st);
    ptr::V3 ro, rd;
    ptr::primary_ray(cam, x, y, width, height, rng, &ro, &rd);
    ptr::Counters c = {0, 0, 0, 0, 0};
    ptr::V3 rad;
    DeviceTex tex;
    bool keep = ptr::trace_sample(sc, k, ro, rd, rng, tex, &rad, &c);
    if (keep && accum) {
        Vec a = accum[i];
        a.x += rad.x; a.y += rad.y; a.z += rad.z;
        accum[i] = a;
    }
    atomicAdd(&totals[0], c.closest_rays);
    atomicAdd(&totals[1], c.shadow_rays);
    atomicAdd(&totals[2], c.nodes_fetched);
    atomicAdd(&totals[3], c.nodes_entered);
    atomicAdd(&totals[4], c.tris_tested);
}

// Arbitrary rays (8 floats each: o.xyz, d.xyz, t_max, unused) against the uploaded scene.
//   closest: the restatement of renderer.cu:371-425 (ptr::closest_hit; its ids are checked against
//            the unmodified kernel's ID plane in every parity test)            -> id, d_min
//   shadow : the reference's OWN, unmodified __device__ function trace_shadow (renderer.cu:273-314),
//            called on the device arrays init_scene_data uploaded                -> 0/1
__global__ void k_trace_rays(const float* rays, int n, int shadow, ptr::SceneView sc, LinearBVHNode* nodes,
                             Object* objs, int* out_id, float* out_t, int* out_occ) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 8;
    if (shadow) {
        Vec o = make_vec(p[0], p[1], p[2]), d = make_vec(p[3], p[4], p[5]);
        out_occ[i] = trace_shadow(o, d, p[6], nodes, objs) ? 1 : 0;
    } else {
        ptr::Counters c = {0, 0, 0, 0, 0};
        float t;
        int id = ptr::closest_hit(sc, ptr::mk(p[0], p[1], p[2]), ptr::mk(p[3], p[4], p[5]), &t, &c);
        if (out_id) out_id[i] = id;
        if (out_t) out_t[i] = t;
    }
}

// host copies of what was last handed to init_scene_data (for the ID-emission trick)
std::vector<Object> g_objects;
std::vector<LinearBVHNode> g_nodes;
std::vector<int> g_lights;
std::vector<std::string> g_tex;
unsigned long long* g_totals = nullptr;

ptr::SceneView scene_view() {
    ptr::SceneView sc;
    sc.objects = reinterpret_cast<const ptr::ObjRec*>(d_objects_ptr);
    sc.nodes = reinterpret_cast<const ptr::NodeRec*>(d_bvh_nodes);
    sc.lights = d_light_indices;
    sc.light_count = d_light_count;
    return sc;
}

int cuda_ok() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[ref] CUDA error: %s\n", cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

std::vector<std::string> split(const char* s) {
    std::vector<std::string> out;
    if (!s) return out;
    std::string cur;
    for (const char* p = s; *p; p++) {
        if (*p == ';') { if (!cur.empty()) out.push_back(cur); cur.clear(); }
        else cur += *p;
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

}  // namespace

extern "C" {

// ---- host surface of the reference --------------------------------------------------
int ref_load_obj(const char* filename, void* out, int cap, const float* offset, float scale,
                 const float* albedo, float metallic, float roughness) {
    std::vector<Object> objs;
    load_obj(filename, objs, Vec{offset[0], offset[1], offset[2]}, scale,
             Vec{albedo[0], albedo[1], albedo[2]}, metallic, roughness);
    if (out) {
        if ((int)objs.size() > cap) return -1;
        memcpy(out, objs.data(), objs.size() * sizeof(Object));
    }
    return (int)objs.size();
}

int ref_bvh_build(void* objects, int n, void* nodes, int cap) {
    std::vector<Object> objs((Object*)objects, (Object*)objects + n);
    BVH bvh;
    bvh.build(objs);
    const std::vector<LinearBVHNode>& nd = bvh.get_nodes();
    if ((int)nd.size() > cap) return -1;
    memcpy(objects, objs.data(), (size_t)n * sizeof(Object));
    memcpy(nodes, nd.data(), nd.size() * sizeof(LinearBVHNode));
    return (int)nd.size();
}

// create_cornell_box() reads "assets/teapot.obj" relative to the working directory.
int ref_create_cornell(void* out, int cap, char* tex_files, int tex_cap) {
    Scene s = create_cornell_box();
    if (out) {
        if ((int)s.objects.size() > cap) return -1;
        memcpy(out, s.objects.data(), s.objects.size() * sizeof(Object));
    }
    if (tex_files && tex_cap > 0) {
        std::string j;
        for (size_t i = 0; i < s.texture_files.size(); i++) { if (i) j += ';'; j += s.texture_files[i]; }
        snprintf(tex_files, tex_cap, "%s", j.c_str());
    }
    return (int)s.objects.size();
}

// CameraController as the reference drives it: default yaw -90 / pitch 0, rotated with
// process_mouse (src/camera.cpp:64-80; sensitivity 0.1 deg per unit).
int ref_camera_params(const float* pos, float mouse_dx, float mouse_dy, int width, int height, void* cam_out) {
    CameraController cam(Vec{pos[0], pos[1], pos[2]}, Vec{0, 0, -1});
    if (mouse_dx != 0.f || mouse_dy != 0.f) cam.process_mouse(mouse_dx, mouse_dy);
    CameraParams p = cam.get_params(width, height);
    memcpy(cam_out, &p, sizeof(p));
    return 0;
}

int ref_tonemap(const void* h_accum, int n, int frames, uint32_t* out) {  // src/pipeline.cpp:59-71
    const Vec* a = (const Vec*)h_accum;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        Vec avg = a[i] * (1.0f / frames);
        int r = toInt(avg.x), g = toInt(avg.y), b = toInt(avg.z);
        out[i] = (255u << 24) | (r << 16) | (g << 8) | b;
    }
    return 0;
}

// ---- renderer boundary ----------------------------------------------------------------
int ref_init_scene(const void* objects, int n, const void* nodes, int n_nodes, const int* lights, int n_lights,
                   const char* tex_files) {
    g_objects.assign((const Object*)objects, (const Object*)objects + n);
    g_nodes.assign((const LinearBVHNode*)nodes, (const LinearBVHNode*)nodes + n_nodes);
    g_lights.assign(lights, lights + n_lights);
    g_tex = split(tex_files);
    init_scene_data(g_objects, g_tex, g_nodes, g_lights);
    if (g_lights.empty()) d_light_indices = nullptr;  // see ref_first_hit_ids
    cudaDeviceSynchronize();
    return cuda_ok();
}

int ref_launch(void* d_accum, int w, int h, int frame_seed, int tx, int ty, const void* cam) {
    CameraParams c;
    memcpy(&c, cam, sizeof(c));
    launch_render_kernel((Vec*)d_accum, w, h, frame_seed, tx, ty, c);
    return 0;
}

// N frames.  cadence 1 = the reference main loop (src/main.cpp:181-192): launch, D2D
// snapshot into d_staging, cudaDeviceSynchronize, every frame.  cadence 0 = launches only.
// Returns elapsed milliseconds (CUDA events on the default stream) in *ms.
int ref_render_frames(void* d_accum, void* d_staging, int w, int h, int first_frame, int n_frames, const void* cam,
                      int cadence, float* ms) {
    CameraParams c;
    memcpy(&c, cam, sizeof(c));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    const size_t bytes = (size_t)w * h * sizeof(Vec);
    for (int f = 0; f < n_frames; f++) {
        launch_render_kernel((Vec*)d_accum, w, h, first_frame + f, 16, 16, c);
        if (cadence) {
            cudaMemcpy(d_staging, d_accum, bytes, cudaMemcpyDeviceToDevice);
            cudaDeviceSynchronize();
        }
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    if (ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cuda_ok();
}

// First-hit ids of one frame from the UNMODIFIED kernel.  h_ids: w*h ints, host.
int ref_first_hit_ids(int w, int h, int frame_seed, const void* cam, int* h_ids) {
    if (g_objects.empty()) return -3;
    if (g_objects.size() >= (1u << 24)) return -1;
    std::vector<Object> coded = g_objects;
    for (size_t k = 0; k < coded.size(); k++) {
        coded[k].emission = Vec{(float)((k & 255) + 1) / 128.f, (float)(((k >> 8) & 255) + 1) / 128.f,
                                (float)(((k >> 16) & 255) + 1) / 128.f};
    }
    std::vector<int> no_lights;
    std::vector<std::string> no_tex;
    init_scene_data(coded, no_tex, g_nodes, no_lights);
    // The reference frees d_light_indices but keeps the stale pointer when the new list is empty
    // (src/renderer.cu:173-183); clear it so the restoring call below does not free it twice.
    d_light_indices = nullptr;
    const size_t n = (size_t)w * h;
    Vec* d_acc = nullptr;
    cudaMalloc(&d_acc, n * sizeof(Vec));
    cudaMemset(d_acc, 0, n * sizeof(Vec));
    CameraParams c;
    memcpy(&c, cam, sizeof(c));
    launch_render_kernel(d_acc, w, h, frame_seed, 16, 16, c);
    std::vector<Vec> host(n);
    cudaMemcpy(host.data(), d_acc, n * sizeof(Vec), cudaMemcpyDeviceToHost);
    cudaFree(d_acc);
    int bad = 0;
    for (size_t i = 0; i < n; i++) {
        const Vec& v = host[i];
        if (v.x == 0.f && v.y == 0.f && v.z == 0.f) { h_ids[i] = -1; continue; }
        float fx = v.x * 128.f - 1.f, fy = v.y * 128.f - 1.f, fz = v.z * 128.f - 1.f;
        int ix = (int)fx, iy = (int)fy, iz = (int)fz;
        if (fx != (float)ix || fy != (float)iy || fz != (float)iz || ix < 0 || iy < 0 || iz < 0 ||
            ix > 255 || iy > 255 || iz > 255) { bad++; h_ids[i] = -2; continue; }
        h_ids[i] = ix | (iy << 8) | (iz << 16);
    }
    // restore the real scene
    init_scene_data(g_objects, g_tex, g_nodes, g_lights);
    if (g_lights.empty()) d_light_indices = nullptr;
    cudaDeviceSynchronize();
    int rc = cuda_ok();
    return rc ? rc : bad;
}

// Instrumented restatement of renderer.cu:319-425 on the currently uploaded scene.
// All outputs are DEVICE pointers (may be NULL).
int ref_primary_counts(int w, int h, int seed_base, int frame_seed, const void* cam, int* d_id, float* d_t,
                       float* d_ray, unsigned* d_fetched, unsigned* d_entered, unsigned* d_tris) {
    ptr::CamRec c;
    memcpy(&c, cam, sizeof(c));
    dim3 threads(16, 16), blocks((w + 15) / 16, (h + 15) / 16);
    k_primary_counts<<<blocks, threads>>>(w, h, seed_base + frame_seed, c, scene_view(), d_id, d_t, d_ray,
                                          d_fetched, d_entered, d_tris);
    cudaDeviceSynchronize();
    return cuda_ok();
}

// Instrumented restatement of the whole sample loop: adds frames first..first+n-1 into
// d_accum (may be NULL) and returns totals[5] = closest rays, shadow rays, nodes fetched,
// nodes entered, triangles tested.
int ref_full_counts(void* d_accum, int w, int h, int seed_base, int first_frame, int n_frames, const void* cam,
                    int max_depth, int rr_threshold, unsigned long long* totals_out) {
    ptr::CamRec c;
    memcpy(&c, cam, sizeof(c));
    if (!g_totals) cudaMalloc(&g_totals, 5 * sizeof(unsigned long long));
    cudaMemset(g_totals, 0, 5 * sizeof(unsigned long long));
    ptr::Consts k = {max_depth, rr_threshold};
    dim3 threads(16, 16), blocks((w + 15) / 16, (h + 15) / 16);
    for (int f = 0; f < n_frames; f++)
        k_full_counts<<<blocks, threads>>>((Vec*)d_accum, w, h, seed_base + first_frame + f, c, scene_view(), k,
                                           g_totals);
    cudaMemcpy(totals_out, g_totals, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return cuda_ok();
}

// Arbitrary-ray oracle (see k_trace_rays).  d_rays / outputs are DEVICE pointers.
int ref_trace_rays(const float* d_rays, int n, int shadow, int* d_id, float* d_t, int* d_occ) {
    if (n <= 0) return 0;
    if (shadow ? !d_occ : !d_id) return -1;
    k_trace_rays<<<(n + 255) / 256, 256>>>(d_rays, n, shadow, scene_view(), d_bvh_nodes, d_objects_ptr, d_id, d_t, d_occ);
    cudaDeviceSynchronize();
    return cuda_ok();
}

int ref_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // extern "C"
