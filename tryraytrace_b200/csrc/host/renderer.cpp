// renderer.cpp -- the reference's two C++ entry points (reference include/renderer.h:35-38,
// :57; src/renderer.cu:134-184, :764-770) as thin wrappers over the C ABI, with a
// process-global context standing in for the reference's file-scope device globals.
#include "renderer.h"
#include "image_io.h"
#include "trt_capi.h"
#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>

namespace {
trt_ctx* g_ctx = nullptr;

trt_ctx* global_ctx() {
    if (!g_ctx) {
        int dev = 0;
        if (const char* e = std::getenv("TRT_DEVICE")) dev = std::atoi(e);
        if (trt_create(dev, &g_ctx) != 0) {
            std::fprintf(stderr, "[Renderer Error] %s\n", trt_last_error());
            g_ctx = nullptr;
        } else {
            // The reference launches on the legacy default stream (src/renderer.cu:769) and its caller relies
            // on that: cudaMemset of the accumulation buffer before, a device-to-device snapshot right after
            // (src/main.cpp:172, :188), both on the legacy stream.  The drop-in entry points keep that order.
            trt_set_stream(g_ctx, (void*)cudaStreamLegacy);
        }
    }
    return g_ctx;
}
}  // namespace

void init_scene_data(const std::vector<Object>& objects, const std::vector<std::string>& texture_files,
                     const std::vector<LinearBVHNode>& nodes, const std::vector<int>& light_indices) {
    trt_ctx* c = global_ctx();
    if (!c) return;
    // Texture slot i belongs to texture_files[i] (objects name it by tex_id).  A file that fails to load keeps
    // its slot -- the reference stores a null handle there (src/renderer.cu:97-126) -- so later indices do not
    // shift: the slot gets a 1x1 placeholder and exactly the objects that named it fall back to untextured.
    static const unsigned char kWhite[3] = {255, 255, 255};
    std::vector<trt_image> imgs;
    std::vector<unsigned char*> owned;
    std::vector<bool> failed;
    for (const std::string& f : texture_files) {
        int w = 0, h = 0;
        unsigned char* rgb = load_ppm(f.c_str(), &w, &h);
        failed.push_back(rgb == nullptr);
        if (rgb) {
            owned.push_back(rgb);
            imgs.push_back(trt_image{w, h, rgb});
        } else {
            imgs.push_back(trt_image{1, 1, kWhite});
        }
    }
    std::vector<Object> objs = objects;
    for (Object& o : objs)
        if (o.tex_id >= (int)imgs.size() || (o.tex_id >= 0 && failed[o.tex_id])) o.tex_id = -1;
    if (trt_upload_scene(c, objs.data(), (int)objs.size(), nodes.data(), (int)nodes.size(), light_indices.data(),
                         (int)light_indices.size(), imgs.data(), (int)imgs.size()) != 0)
        std::fprintf(stderr, "[Renderer Error] %s\n", trt_last_error());
    else
        std::printf("[Renderer] Uploaded %zu objects, %zu BVH nodes, %zu lights.\n", objects.size(), nodes.size(),
                    light_indices.size());
    for (unsigned char* p : owned) std::free(p);
}

void launch_render_frames(Vec* accum_buffer, int width, int height, int first_frame_seed, int n_frames,
                          CameraParams cam) {
    trt_ctx* c = global_ctx();
    if (!c) return;
    if (trt_render(c, reinterpret_cast<float*>(accum_buffer), width, height, first_frame_seed, n_frames, 1, &cam,
                   nullptr) != 0)
        std::fprintf(stderr, "[Renderer Error] %s\n", trt_last_error());
}

void launch_render_kernel(Vec* accum_buffer, int width, int height, int frame_seed, int tx, int ty,
                          CameraParams cam) {
    (void)tx;
    (void)ty;
    launch_render_frames(accum_buffer, width, height, frame_seed, 1, cam);
}
