// loader.cpp -- OBJ reader behind load_obj (reference src/loader.cpp:22-103).
//
// Behaviour kept from the reference, because BVH order and therefore every object
// id depends on it:
//   * 256-byte fgets lines; only "v x y z" and "f a b c" records are read (:37-94);
//   * a vertex is stored as v*scale+offset (:51).  The reference is built with
//     -march=native, where g++ contracts that expression into one FMA per
//     component; fmaf() pins the same rounding whatever flags this file gets;
//   * faces are 1-based bare indices; a record sscanf cannot read three ints from
//     (e.g. "f 1/1 2/2 3/3") is dropped, extra indices after the third are ignored,
//     out-of-range indices drop the face (:63-75);
//   * material: albedo/metallic/roughness from the arguments, tex_id -1, every other
//     field zero (:84-92).
#include "loader.h"
#include <cmath>
#include <cstdio>
#include <cstring>

void load_obj(const char* filename, std::vector<Object>& objects,
              Vec offset, float scale, Vec albedo, float metallic, float roughness) {
    FILE* fp = std::fopen(filename, "r");
    if (!fp) {
        std::printf("[Loader Error] Cannot open file: %s\n", filename);
        return;
    }
    std::vector<Vec> verts;
    verts.reserve(4096);

    Object proto;
    std::memset(&proto, 0, sizeof(proto));
    proto.albedo.x = albedo.x;
    proto.albedo.y = albedo.y;
    proto.albedo.z = albedo.z;
    proto.metallic = metallic;
    proto.roughness = roughness;
    proto.tex_id = -1;

    char buf[256];
    while (std::fgets(buf, sizeof(buf), fp)) {
        if (buf[1] != ' ') continue;
        if (buf[0] == 'v') {
            Vec p{0.f, 0.f, 0.f};
            std::sscanf(buf, "v %f %f %f", &p.x, &p.y, &p.z);
            Vec q;
            std::memset(&q, 0, sizeof(q));  // keep the pad lane deterministic
            q.x = std::fmaf(p.x, scale, offset.x);
            q.y = std::fmaf(p.y, scale, offset.y);
            q.z = std::fmaf(p.z, scale, offset.z);
            verts.push_back(q);
        } else if (buf[0] == 'f') {
            int a, b, c;
            if (std::sscanf(buf, "f %d %d %d", &a, &b, &c) != 3) continue;
            const int n = (int)verts.size();
            if (a < 1 || b < 1 || c < 1 || a > n || b > n || c > n) continue;
            Object o;
            std::memcpy(&o, &proto, sizeof(o));
            std::memcpy(&o.v0, &verts[a - 1], sizeof(Vec));
            std::memcpy(&o.v1, &verts[b - 1], sizeof(Vec));
            std::memcpy(&o.v2, &verts[c - 1], sizeof(Vec));
            objects.push_back(o);
        }
    }
    std::fclose(fp);
    std::printf("[Loader] Loaded: %s\n         Vertices: %lu\n", filename, (unsigned long)verts.size());
}
