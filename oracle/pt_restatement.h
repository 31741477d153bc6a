// pt_restatement.h -- TEST INFRASTRUCTURE (oracle).  Not part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may
// use anything under oracle/.
//
// A restatement of the reference path tracer's per-sample algorithm
// (/root/reference/src/renderer.cu:317-760 and the helpers it calls) as one
// host+device template, so that the same text can be
//   * compiled by g++ into the CPU oracle (oracle/cpu_oracle.cpp), and
//   * compiled by nvcc, with the reference's own flags, into an instrumented GPU
//     kernel next to the unmodified reference kernel (oracle/ref_gpu.cu) to count
//     rays / node visits, which the unmodified kernel cannot report.
//
// Every function cites the reference lines it follows.  Expression shapes are kept
// the same as the reference's so that nvcc contracts the same multiply-adds.
// The CPU build cannot be bit-identical to a GPU run (MUFU approximations, FTZ); the
// parity oracle proper is the unmodified reference kernel in oracle/_ref.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ inline
#else
#define PT_HD inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ptr {

// ---- include/common.h:24-97 -------------------------------------------------------
struct V3 {
    float x, y, z, _pad;
};
PT_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; v._pad = 0.f; return v; }
PT_HD V3 operator+(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD V3 operator-(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD V3 operator*(const V3& a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
PT_HD V3 mult(const V3& a, const V3& b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD V3 cross(const V3& a, const V3& b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PT_HD float length(const V3& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
PT_HD V3 normalized(V3 a) {  // Vec::norm, common.h:70-78
    float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 0) {
        float inv = 1.0f / len;
        a.x *= inv; a.y *= inv; a.z *= inv;
    }
    return a;
}

// ---- record layouts (include/scene.h:30-55, :64-72; include/bvh.h:12-28) -------------
struct ObjRec {  // 112 bytes
    V3 v0, v1, v2, albedo, emission;
    float metallic, roughness, ior, transmission;
    int tex_id;
    float pad1, pad2, pad3;
};
struct NodeRec {  // 48 bytes
    V3 bmin, bmax;
    int a;  // left child / primitive offset
    int b;  // right child / primitive count
    int axis, is_leaf;
};
struct CamRec {  // 80 bytes
    V3 pos, cx, cy, dir;
    float lens_radius, focus_dist;
    float _pad[2];
};

struct SceneView {
    const ObjRec* objects;
    const NodeRec* nodes;
    const int* lights;
    int light_count;
};

struct Counters {
    unsigned long long closest_rays, shadow_rays, nodes_fetched, nodes_entered, tris_tested;
};

struct Consts {
    int max_depth;     // 30, renderer.cu:363
    int rr_threshold;  // 3,  renderer.cu:364
};

// ---- include/aabb.h:5-6, :49-69 -----------------------------------------------------
PT_HD float fmin_w(float a, float b) { return a < b ? a : b; }
PT_HD float fmax_w(float a, float b) { return a > b ? a : b; }

PT_HD bool slab_hit(const NodeRec& n, const V3& o, const V3& inv, float t_min, float t_max) {
    float tx1 = (n.bmin.x - o.x) * inv.x;
    float tx2 = (n.bmax.x - o.x) * inv.x;
    float tmin = fmin_w(tx1, tx2);
    float tmax = fmax_w(tx1, tx2);
    float ty1 = (n.bmin.y - o.y) * inv.y;
    float ty2 = (n.bmax.y - o.y) * inv.y;
    tmin = fmax_w(tmin, fmin_w(ty1, ty2));
    tmax = fmin_w(tmax, fmax_w(ty1, ty2));
    float tz1 = (n.bmin.z - o.z) * inv.z;
    float tz2 = (n.bmax.z - o.z) * inv.z;
    tmin = fmax_w(tmin, fmin_w(tz1, tz2));
    tmax = fmin_w(tmax, fmax_w(tz1, tz2));
    return tmax >= tmin && tmax > t_min && tmin < t_max;
}

// ---- renderer.cu:235-268  Moeller-Trumbore, 0 = miss --------------------------------
PT_HD float tri_hit(const ObjRec& obj, const V3& r_o, const V3& r_d) {
    const float eps = 1e-5f;
    V3 e1 = obj.v1 - obj.v0;
    V3 e2 = obj.v2 - obj.v0;
    V3 h = cross(r_d, e2);
    float a = dot(e1, h);
    if (a > -eps && a < eps) return 0.0f;
    float f = 1.0f / a;
    V3 s = r_o - obj.v0;
    float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return 0.0f;
    V3 q = cross(s, e1);
    float v = f * dot(r_d, q);
    if (v < 0.0f || u + v > 1.0f) return 0.0f;
    float t = f * dot(e2, q);
    if (t > eps) return t;
    return 0.0f;
}

// ---- renderer.cu:371-425  closest hit, reference order ------------------------------
PT_HD V3 safe_inverse(const V3& d) {  // :371-379
    V3 r;
    r.x = (fabsf(d.x) < 1e-8f) ? (d.x >= 0 ? 1e20f : -1e20f) : (1.0f / d.x);
    r.y = (fabsf(d.y) < 1e-8f) ? (d.y >= 0 ? 1e20f : -1e20f) : (1.0f / d.y);
    r.z = (fabsf(d.z) < 1e-8f) ? (d.z >= 0 ? 1e20f : -1e20f) : (1.0f / d.z);
    r._pad = 0.f;
    return r;
}

PT_HD int closest_hit(const SceneView& sc, const V3& r_o, const V3& r_d, float* t_out, Counters* cnt) {
    V3 inv = safe_inverse(r_d);
    float d_min = 1e20f;
    int id = -1;
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    if (cnt) cnt->closest_rays++;
    while (sp > 0) {
        int ni = stack[--sp];
        NodeRec node = sc.nodes[ni];
        if (cnt) cnt->nodes_fetched++;
        if (!slab_hit(node, r_o, inv, 0.0f, d_min)) continue;
        if (cnt) cnt->nodes_entered++;
        if (node.is_leaf) {
            for (int k = 0; k < node.b; k++) {
                int oi = node.a + k;
                if (cnt) cnt->tris_tested++;
                float t = tri_hit(sc.objects[oi], r_o, r_d);
                if (t > 0.0f && t < d_min) {
                    d_min = t;
                    id = oi;
                }
            }
        } else {
            stack[sp++] = node.b;  // right first, so left is popped first (:422-423)
            stack[sp++] = node.a;
        }
    }
    *t_out = d_min;
    return id;
}

// ---- renderer.cu:273-314  any hit ------------------------------------------------------
PT_HD bool shadow_hit(const SceneView& sc, const V3& origin, const V3& dir, float max_dist, Counters* cnt) {
    V3 inv = mk(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    if (cnt) cnt->shadow_rays++;
    while (sp > 0) {
        int ni = stack[--sp];
        NodeRec node = sc.nodes[ni];
        if (cnt) cnt->nodes_fetched++;
        if (!slab_hit(node, origin, inv, 0.001f, max_dist)) continue;
        if (cnt) cnt->nodes_entered++;
        if (node.is_leaf) {
            for (int k = 0; k < node.b; k++) {
                if (cnt) cnt->tris_tested++;
                float t = tri_hit(sc.objects[node.a + k], origin, dir);
                if (t > 0.001f && t < max_dist - 0.001f) return true;
            }
        } else {
            stack[sp++] = node.b;
            stack[sp++] = node.a;
        }
    }
    return false;
}

// ---- renderer.cu:188-192, :201-204, :207-227 -----------------------------------------
PT_HD float tri_area(const ObjRec& o) {
    V3 e1 = o.v1 - o.v0;
    V3 e2 = o.v2 - o.v0;
    return length(cross(e1, e2)) * 0.5f;
}
PT_HD V3 schlick(float cosine, V3 F0) {
    return F0 + (mk(1.0f, 1.0f, 1.0f) - F0) * powf(1.0f - cosine, 5.0f);
}
template <class Rng>
PT_HD V3 rough_reflection(V3 perfect, float roughness, Rng& rng) {
    float r1 = rng.uniform() * 2.0f * M_PI;
    float r2 = rng.uniform();
    float z = 1.0f - 2.0f * r2;
    float r = sqrtf(1.0f - z * z);
    V3 sph = mk(r * cosf(r1), r * sinf(r1), z);
    return normalized(perfect + sph * roughness);
}

// ---- renderer.cu:331-356  primary ray ---------------------------------------------------
template <class Rng>
PT_HD void primary_ray(const CamRec& cam, int x, int y, int width, int height, Rng& rng, V3* ro, V3* rd) {
    float r1 = 2 * rng.uniform();
    float r2 = 2 * rng.uniform();
    float dx = r1 < 1 ? sqrtf(r1) - 1 : 1 - sqrtf(2 - r1);
    float dy = r2 < 1 ? sqrtf(r2) - 1 : 1 - sqrtf(2 - r2);
    V3 dir_pinhole = normalized(cam.cx * (((x + .5f + dx) / width - .5f)) +
                                cam.cy * (((y + .5f + dy) / height - .5f)) + cam.dir);
    V3 lens_offset = mk(0, 0, 0);
    if (cam.lens_radius > 0.0f) {
        float lr = cam.lens_radius * sqrtf(rng.uniform());
        float ltheta = 2 * M_PI * rng.uniform();
        V3 u = normalized(cam.cx);
        V3 v = normalized(cam.cy);
        lens_offset = u * (lr * cosf(ltheta)) + v * (lr * sinf(ltheta));
    }
    V3 p_focus = cam.pos + dir_pinhole * cam.focus_dist;
    V3 r_o = cam.pos + lens_offset;
    *ro = r_o;
    *rd = normalized(p_focus - r_o);
}

// ---- renderer.cu:359-756  one sample; returns false when the sample is dropped (:739-742)
// Tex: functor  V3 operator()(int tex_id, float u, float v)  (tex2D at :478).
template <class Rng, class Tex>
PT_HD bool trace_sample(const SceneView& sc, const Consts& k, V3 r_o, V3 r_d, Rng& rng, const Tex& tex,
                        V3* out, Counters* cnt) {
    V3 throughput = mk(1, 1, 1);
    V3 radiance = mk(0, 0, 0);
    int prev_mode = 1;  // SPEC (:365); 0 DIFF, 1 SPEC, 2 REFR

    for (int depth = 0; depth < k.max_depth; depth++) {
        float d_min;
        int id = closest_hit(sc, r_o, r_d, &d_min, cnt);
        if (id < 0) break;  // :427

        const ObjRec& obj = sc.objects[id];
        V3 x_hit = r_o + r_d * d_min;
        V3 e1 = obj.v1 - obj.v0;
        V3 e2 = obj.v2 - obj.v0;
        V3 n = normalized(cross(e1, e2));
        V3 nl = dot(n, r_d) < 0 ? n : n * -1;

        V3 albedo = obj.albedo;
        float metallic = obj.metallic;
        float roughness = obj.roughness;
        float transmission = obj.transmission;

        if (obj.tex_id >= 0) {  // :465-481 planar mapping
            const float scale = 0.01f;
            float u, v;
            if (fabsf(n.y) > 0.9f)      { u = x_hit.x; v = x_hit.z; }
            else if (fabsf(n.x) > 0.9f) { u = x_hit.z; v = x_hit.y; }
            else                        { u = x_hit.x; v = x_hit.y; }
            u *= scale; v *= scale;
            v = 1.0f - v;
            albedo = mult(albedo, tex(obj.tex_id, u, v));
        }

        if (prev_mode == 1 || prev_mode == 2)  // :489-495
            radiance = radiance + mult(throughput, obj.emission);
        if (obj.emission.x > 0.001f || obj.emission.y > 0.001f || obj.emission.z > 0.001f) break;  // :497-499

        // :509-556 lobe weights
        float diffuse_suppression = powf(1.0f - metallic, 2.0f);
        float spec_attenuation = 1.0f - (roughness * roughness);
        if (spec_attenuation < 0.0f) spec_attenuation = 0.0f;
        V3 F0 = mk(0.04f, 0.04f, 0.04f);
        F0 = F0 * (1.0f - metallic) + albedo * metallic;
        float cos_theta = fmaxf(dot(nl, r_d * -1.0f), 0.0f);
        V3 F = schlick(cos_theta, F0);
        float F_avg = (F.x + F.y + F.z) / 3.0f;
        float w_spec = F_avg * spec_attenuation;
        float w_trans = (1.0f - F_avg) * transmission;
        float albedo_lum = fmaxf(albedo.x, fmaxf(albedo.y, albedo.z));
        float w_diff = (1.0f - F_avg) * (1.0f - transmission) * diffuse_suppression * albedo_lum;
        float sum = w_spec + w_trans + w_diff;
        if (sum < 1e-5f) { w_diff = 1.0f; sum = 1.0f; }
        float p_spec = w_spec / sum;
        float p_trans = w_trans / sum;

        if (depth > k.rr_threshold) {  // :559-565
            float p = fmaxf(albedo.x, fmaxf(albedo.y, albedo.z));
            if (p < 0.05f) p = 0.05f;
            if (rng.uniform() < p) throughput = throughput * (1.0f / p);
            else break;
        }

        float rnd = rng.uniform();  // :567

        if (rnd < p_spec) {  // :571-589
            V3 perfect = r_d - n * 2 * dot(n, r_d);
            r_d = rough_reflection(perfect, roughness, rng);
            if (dot(r_d, nl) <= 0.0f) break;
            float weight = 1.0f / p_spec;
            throughput = mult(throughput, F) * weight;
            r_o = x_hit + nl * 1e-3f;
            prev_mode = 1;
        } else if (rnd < p_spec + p_trans) {  // :592-648
            bool into = dot(n, nl) > 0;
            float nc = 1.0f;
            float nt = obj.ior;
            float nnt = into ? nc / nt : nt / nc;
            float ddn = dot(r_d, nl);
            float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
            if (cos2t < 0.0f) {
                V3 perfect = r_d - n * 2.0f * dot(n, r_d);
                r_d = rough_reflection(perfect, roughness, rng);
                r_o = x_hit + r_d * 1e-4f;
            } else {
                V3 tdir = normalized(r_d * nnt - n * ((into ? 1.0f : -1.0f) * (ddn * nnt + sqrtf(cos2t))));
                if (roughness > 0.0f) {
                    float r1 = rng.uniform() * 2.0f * M_PI;
                    float r2 = rng.uniform();
                    float z = 1.0f - 2.0f * r2;
                    float r = sqrtf(1.0f - z * z);
                    V3 rv = mk(r * cosf(r1), r * sinf(r1), z);
                    tdir = normalized(tdir + rv * roughness);
                }
                r_d = tdir;
                r_o = x_hit + r_d * 1e-4f;
            }
            float p_branch = (1.0f - p_spec) * transmission;
            if (p_branch > 1e-4f) throughput = mult(throughput, albedo) * (1.0f / p_branch);
            prev_mode = 2;
        } else {  // :651-733 diffuse + next-event estimation
            if (sc.light_count > 0) {
                int l_idx = (int)(rng.uniform() * (sc.light_count - 0.001f));
                const ObjRec& light = sc.objects[sc.lights[l_idx]];
                float r1 = rng.uniform();
                float r2 = rng.uniform();
                float sqr1 = sqrtf(r1);
                float u = 1.0f - sqr1;
                float v = sqr1 * (1.0f - r2);
                V3 light_pos = light.v0 * u + light.v1 * v + light.v2 * (1.0f - u - v);
                V3 to_light = light_pos - x_hit;
                float dist_sq = dot(to_light, to_light);
                if (dist_sq < 5) dist_sq = 5;
                float dist = sqrtf(dist_sq);
                V3 L_dir = to_light * (1.0f / dist);
                float cos_t = dot(nl, L_dir);
                V3 light_n = normalized(cross(light.v1 - light.v0, light.v2 - light.v0));
                float cos_light = -dot(light_n, L_dir);
                if (cos_t > 0.0f && cos_light > 0.0f) {
                    if (!shadow_hit(sc, x_hit + nl * 1e-3f, L_dir, dist - 1e-2f, cnt)) {
                        float area = tri_area(light);
                        float pdf = 1.0f / (area * sc.light_count);
                        float G = (cos_t * cos_light) / dist_sq;
                        V3 brdf = albedo * (1.0f / M_PI);
                        V3 contribution = mult(light.emission, brdf) * (G / pdf);
                        radiance = radiance + mult(throughput, contribution);
                    }
                }
            }
            V3 diffuse = albedo * (1.0f - metallic);
            float r1 = 2 * M_PI * rng.uniform();
            float r2 = rng.uniform();
            float r2s = sqrtf(r2);
            V3 w = nl;
            V3 temp = (fabs(w.x) > 0.1f ? mk(0, 1, 0) : mk(1, 0, 0));
            V3 u = normalized(cross(temp, w));
            V3 v = cross(w, u);
            r_d = normalized(u * cosf(r1) * r2s + v * sinf(r1) * r2s + w * sqrtf(1 - r2));
            float p_diff = 1.0f - p_spec - (1.0f - p_spec) * transmission;
            float weight = 1.0f / p_diff;
            throughput = mult(throughput, diffuse) * weight;
            r_o = x_hit + nl * 1e-3f;
            prev_mode = 0;
        }
    }

    // :739-756 sample filter
    if (isnan(radiance.x) || isnan(radiance.y) || isnan(radiance.z) ||
        isinf(radiance.x) || isinf(radiance.y) || isinf(radiance.z))
        return false;
    if (radiance.x < 0.0f) radiance.x = 0.0f;
    if (radiance.y < 0.0f) radiance.y = 0.0f;
    if (radiance.z < 0.0f) radiance.z = 0.0f;
    float max_lum = 100.0f;
    float lum = radiance.x * 0.21 + radiance.y * 0.71 + radiance.z * 0.07;
    if (lum > max_lum) radiance = radiance * (max_lum / lum);
    *out = radiance;
    return true;
}

}  // namespace ptr
