// scene.h -- scene data model at the renderer boundary.
// Drop-in for the reference's include/scene.h: Refl_t (:13-17), Object (:30-55,
// one triangle + material, 112 bytes), CameraParams (:64-72, 80 bytes), Scene
// (:80-85), create_cornell_box (:88).  Offsets are pinned by static_asserts
// against SURVEY Appendix B.1.
#pragma once
#include "common.h"
#include "aabb.h"
#include <cstddef>
#include <string>
#include <vector>

enum Refl_t { DIFF, SPEC, REFR };

struct Object {
    Vec v0, v1, v2;      // triangle vertices (world space, scale/offset baked in)
    Vec albedo;          // base colour
    Vec emission;        // radiance emitted (light sources)
    float metallic;      // 0 dielectric .. 1 metal
    float roughness;     // 0 mirror .. 1 diffuse
    float ior;           // index of refraction
    float transmission;  // 0 opaque .. 1 glass
    int tex_id;          // index into Scene::texture_files, -1 = untextured
    float pad1, pad2, pad3;
};

static_assert(sizeof(Object) == 112, "Object layout");
static_assert(offsetof(Object, albedo) == 48 && offsetof(Object, emission) == 64, "Object layout");
static_assert(offsetof(Object, metallic) == 80 && offsetof(Object, tex_id) == 96, "Object layout");

struct CameraParams {
    Vec pos;   // eye
    Vec cx;    // image-plane x axis, fov and aspect folded in
    Vec cy;    // image-plane y axis, fov folded in
    Vec dir;   // unit view direction
    float lens_radius;  // aperture / 2, 0 = pinhole
    float focus_dist;
};

static_assert(sizeof(CameraParams) == 80 && offsetof(CameraParams, lens_radius) == 64, "CameraParams layout");

struct Scene {
    std::vector<Object> objects;
    std::vector<std::string> texture_files;
    AABB world_bound;
};

// The reference's stock scene (src/scene.cpp:24-123).
Scene create_cornell_box();

// Benchmark scenes C1..C5 of SURVEY section 8(d) ("config" = 1..5).  `asset_dir`
// holds the OBJ/PPM inputs; `grid` overrides the C5 instancing grid (0 = 40).
Scene create_config_scene(int config, const char* asset_dir, int grid = 0);
