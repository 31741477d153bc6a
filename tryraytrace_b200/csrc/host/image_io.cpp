// image_io.cpp -- PPM in/out: save_snapshot (reference src/image_io.cpp:17-92) and
// the P6 texture reader (reference src/renderer.cu:36-76).
#include "image_io.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include <sys/stat.h>

void save_snapshot(const Vec* h_accum, int w, int h, int frame, float focus_dist, float aperture) {
    mkdir("logs", 0755);
    char stamp[64];
    std::time_t now = std::time(nullptr);
    std::strftime(stamp, sizeof(stamp), "%Y-%m-%d_%H-%M-%S", std::localtime(&now));
    char path[256];
    std::snprintf(path, sizeof(path), "logs/%s_Frame%d_F%.1f_A%.2f.ppm", stamp, frame, focus_dist, aperture);

    const size_t n = (size_t)w * h;
    std::vector<unsigned char> rgb(n * 3);
    const float inv = 1.0f / frame;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        const Vec a = h_accum[i] * inv;
        rgb[i * 3 + 0] = (unsigned char)toInt(a.x);
        rgb[i * 3 + 1] = (unsigned char)toInt(a.y);
        rgb[i * 3 + 2] = (unsigned char)toInt(a.z);
    }
    FILE* fp = std::fopen(path, "wb");
    if (!fp) {
        std::fprintf(stderr, "[IO Error] Failed to open file for writing: %s\n", path);
        return;
    }
    std::fprintf(fp, "P6\n%d %d\n%d\n", w, h, 255);
    std::fwrite(rgb.data(), 1, rgb.size(), fp);
    std::fclose(fp);
    std::printf("[IO] Snapshot saved: %s\n", path);
}

unsigned char* load_ppm(const char* filename, int* w, int* h) {
    FILE* fp = std::fopen(filename, "rb");
    if (!fp) {
        std::printf("[Texture Error] Cannot open file: %s\n", filename);
        return nullptr;
    }
    char magic[64] = {0};
    int maxval = 0;
    if (std::fscanf(fp, "%63s", magic) != 1 || std::strcmp(magic, "P6") != 0) {
        std::printf("[Texture Error] Not a P6 binary PPM: %s\n", filename);
        std::fclose(fp);
        return nullptr;
    }
    if (std::fscanf(fp, "%d %d %d", w, h, &maxval) != 3 || *w <= 0 || *h <= 0) {
        std::printf("[Texture Error] Bad header: %s\n", filename);
        std::fclose(fp);
        return nullptr;
    }
    std::fgetc(fp);  // the single whitespace byte that ends the header
    const size_t bytes = (size_t)(*w) * (*h) * 3;
    unsigned char* data = (unsigned char*)std::malloc(bytes);
    if (!data || std::fread(data, 1, bytes, fp) != bytes) {
        std::printf("[Texture Error] Unexpected EOF: %s\n", filename);
        std::free(data);
        std::fclose(fp);
        return nullptr;
    }
    std::fclose(fp);
    std::printf("[Texture] Loaded: %s (%dx%d)\n", filename, *w, *h);
    return data;
}
