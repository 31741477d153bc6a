// scene.cpp -- scene factories.
//
// create_cornell_box() rebuilds the reference's stock scene (src/scene.cpp:24-123):
// a room of seven single-triangle surfaces (floor, ceiling, textured back wall,
// black mirror behind the camera, red and green side walls, a small ceiling light)
// plus a metallic teapot.  create_config_scene() builds the benchmark scenes C1..C5
// of SURVEY section 8(d) from the same room shell and the OBJ assets.
#include "scene.h"
#include "loader.h"
#include <cstdio>
#include <cstring>
#include <string>

namespace {

struct Surface {
    float v[9];
    float albedo[3];
    float emission;
    float metallic, roughness, ior;
    int tex_id;
};

// Room shell, in the order the reference pushes it (src/scene.cpp:59-91).
const Surface kShell[7] = {
    {{-50, 0, 0, 50, 0, 600, 150, 0, 0},        {0.75f, 0.75f, 0.75f}, 0.f, 0.f, 1.f, 1.45f, -1},  // floor
    {{-50, 100, 0, 150, 100, 0, 50, 100, 600},  {0.75f, 0.75f, 0.75f}, 0.f, 0.f, 1.f, 1.45f, -1},  // ceiling
    {{-50, 0, 0, 150, 0, 0, 50, 200, 0},        {0.75f, 0.75f, 0.75f}, 0.f, 0.f, 1.f, 1.45f, 0},   // back wall (textured)
    {{-50, 0, 300, 150, 0, 300, 50, 200, 300},  {0.f, 0.f, 0.f},       0.f, 1.f, 0.f, 0.f,   -1},  // mirror at z=300
    {{0, 0, -50, 0, 200, 50, 0, 0, 550},        {0.75f, 0.25f, 0.25f}, 0.f, 0.f, 1.f, 1.45f, -1},  // left, red
    {{100, 0, 550, 100, 200, 50, 100, 0, -50},  {0.25f, 0.75f, 0.25f}, 0.f, 0.f, 1.f, 1.45f, -1},  // right, green
    {{30, 99.9f, 30, 70, 99.9f, 30, 50, 99.9f, 50}, {0.f, 0.f, 0.f},  20.f, 0.f, 1.f, 1.45f, -1},  // light
};

Object make_object(const Surface& s) {
    Object o;
    std::memset(&o, 0, sizeof(o));
    o.v0 = Vec{s.v[0], s.v[1], s.v[2]};
    o.v1 = Vec{s.v[3], s.v[4], s.v[5]};
    o.v2 = Vec{s.v[6], s.v[7], s.v[8]};
    o.albedo = Vec{s.albedo[0], s.albedo[1], s.albedo[2]};
    o.emission = Vec{s.emission, s.emission, s.emission};
    o.metallic = s.metallic;
    o.roughness = s.roughness;
    o.ior = s.ior;
    o.transmission = 0.f;
    o.tex_id = s.tex_id;
    return o;
}

void add_shell(Scene& scene, bool textured_back_wall) {
    for (const Surface& s : kShell) {
        Object o = make_object(s);
        if (!textured_back_wall) o.tex_id = -1;
        scene.objects.push_back(o);
    }
}

void finish_bounds(Scene& scene) {
    AABB b = AABB::empty();
    for (const Object& o : scene.objects) {
        b.grow(o.v0);
        b.grow(o.v1);
        b.grow(o.v2);
    }
    b.min = b.min - make_vec(0.1f, 0.1f, 0.1f);
    b.max = b.max + make_vec(0.1f, 0.1f, 0.1f);
    scene.world_bound = b;
}

std::string asset(const char* dir, const char* name) {
    std::string p = dir ? dir : "assets";
    if (!p.empty() && p.back() != '/') p += '/';
    return p + name;
}

const Vec kWhite{0.75f, 0.75f, 0.75f};

}  // namespace

Scene create_cornell_box() {
    Scene scene;
    scene.texture_files.push_back("assets/earth.ppm");
    add_shell(scene, true);
    // the stock teapot: metallic 1, roughness 0.1 -> always the specular lobe
    load_obj("assets/teapot.obj", scene.objects, Vec{50.0f, 10.0f, 50.0f}, 10.0f, kWhite, 1.0f, 0.1f);
    std::printf("[Scene] Scene created with %lu objects.\n", (unsigned long)scene.objects.size());
    finish_bounds(scene);
    std::printf("[Scene] World Bound: Min(%.1f, %.1f, %.1f) Max(%.1f, %.1f, %.1f)\n",
                scene.world_bound.min.x, scene.world_bound.min.y, scene.world_bound.min.z,
                scene.world_bound.max.x, scene.world_bound.max.y, scene.world_bound.max.z);
    return scene;
}

Scene create_config_scene(int config, const char* asset_dir, int grid) {
    Scene scene;
    switch (config) {
    case 0: {
        // stock scene with a caller-supplied asset directory
        scene.texture_files.push_back(asset(asset_dir, "earth.ppm"));
        add_shell(scene, true);
        load_obj(asset(asset_dir, "teapot.obj").c_str(), scene.objects, Vec{50.0f, 10.0f, 50.0f}, 10.0f,
                 kWhite, 1.0f, 0.1f);
        break;
    }
    case 1:  // C1: room + cube, 19 triangles
        add_shell(scene, false);
        load_obj(asset(asset_dir, "cube.obj").c_str(), scene.objects, Vec{50.f, 15.f, 60.f}, 15.f, kWhite, 0.f, 1.f);
        break;
    case 2:  // C2: room + teapot, 6327 triangles (the headline scene)
        add_shell(scene, false);
        load_obj(asset(asset_dir, "teapot.obj").c_str(), scene.objects, Vec{48.f, 5.f, 80.f}, 14.f, kWhite, 0.f, 1.f);
        break;
    case 3:  // C3: textured back wall + cow + teddy, 9003 triangles
        scene.texture_files.push_back(asset(asset_dir, "earth.ppm"));
        add_shell(scene, true);
        load_obj(asset(asset_dir, "cow.obj").c_str(), scene.objects, Vec{32.f, 19.f, 70.f}, 5.f, kWhite, 0.f, 1.f);
        load_obj(asset(asset_dir, "teddy.obj").c_str(), scene.objects, Vec{72.f, 19.f, 95.f}, 0.9f, kWhite, 0.f, 1.f);
        break;
    case 4:  // C4: room + pumpkin, 10007 triangles
        add_shell(scene, false);
        load_obj(asset(asset_dir, "pumpkin.obj").c_str(), scene.objects, Vec{51.6f, 29.5f, 146.f}, 0.6f, kWhite, 0.f, 1.f);
        break;
    case 5: {  // C5: floor + light + grid x grid teapots
        const int g = grid > 0 ? grid : 40;
        Surface fl = {{-4000, 0, -4000, 0, 0, 8000, 4000, 0, -4000}, {0.75f, 0.75f, 0.75f}, 0.f, 0.f, 1.f, 1.45f, -1};
        Surface li = {{-2000, 800, -2000, 2000, 800, -2000, 0, 800, 2000}, {0.f, 0.f, 0.f}, 20.f, 0.f, 1.f, 1.45f, -1};
        scene.objects.push_back(make_object(fl));
        scene.objects.push_back(make_object(li));
        // parse the mesh once, then instance it (the reference loader would re-read the file per instance)
        std::vector<Object> unit;
        load_obj(asset(asset_dir, "teapot.obj").c_str(), unit, Vec{0.f, 0.f, 0.f}, 1.0f, kWhite, 0.f, 1.f);
        scene.objects.reserve(2 + (size_t)unit.size() * g * g);
        for (int iz = 0; iz < g; iz++)
            for (int ix = 0; ix < g; ix++) {
                const Vec off{-175.f + 9.f * ix, 0.f, 60.f - 9.f * iz};
                for (const Object& u : unit) {
                    Object o = u;
                    const Vec* src[3] = {&u.v0, &u.v1, &u.v2};
                    Vec* dst[3] = {&o.v0, &o.v1, &o.v2};
                    for (int k = 0; k < 3; k++)
                        *dst[k] = Vec{std::fmaf(src[k]->x, 1.2f, off.x), std::fmaf(src[k]->y, 1.2f, off.y),
                                      std::fmaf(src[k]->z, 1.2f, off.z)};
                    scene.objects.push_back(o);
                }
            }
        break;
    }
    default:
        std::printf("[Scene Error] unknown config %d\n", config);
        return scene;
    }
    finish_bounds(scene);
    return scene;
}
