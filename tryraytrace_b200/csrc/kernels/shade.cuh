// shade.cuh -- material evaluation for one path vertex: hit record, planar texture,
// emission rule, lobe selection, Russian roulette, next-event estimation set-up and the
// three bounce types (reference src/renderer.cu:434-733, helpers :188-227).
//
// The estimator, including its quirks, is reproduced as is: emission
// only after a specular/transmissive vertex, dist^2 floored at 5 with an un-normalised
// light direction below that, p_diff computed from p_spec and transmission, double
// precision pi.  Arithmetic here is ordinary fast-math FP32 -- radiance parity is a
// tolerance gate, not a bit-exact one; the decisions that must be bit-exact (which
// triangle a ray hits) are made in the traversal code from the ray this function writes.
#pragma once
#include "common.cuh"
#include "xorwow.cuh"
#include "raygen.cuh"

namespace trt {

TRT_DEV F3 v_add(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
TRT_DEV F3 v_sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
TRT_DEV F3 v_scale(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
TRT_DEV F3 v_mul(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
TRT_DEV float v_dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TRT_DEV F3 v_cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
TRT_DEV F3 v_norm(F3 a) {
    const float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 0.f) {
        const float r = 1.0f / len;
        a = f3(a.x * r, a.y * r, a.z * r);
    }
    return a;
}
TRT_DEV F3 ld3(const float4 v) { return f3(v.x, v.y, v.z); }

enum { MODE_DIFF = 0, MODE_SPEC = 1, MODE_REFR = 2 };

struct PathVertexIO {
    // in/out
    Ray ray;
    F3 thr, rad;
    int depth;      // bounce index of this vertex (0 = camera ray's hit)
    int prev_mode;  // MODE_* of the previous bounce (SPEC for the camera ray)
    Xorwow rng;
    // out: shadow-ray request for the next-event estimate
    bool shadow;
    Ray shadow_ray;
    float shadow_max_dist;
    F3 shadow_contrib;  // already multiplied by the path throughput
};

// uniform-sphere perturbation of a mirror direction (reference :207-227)
TRT_DEV F3 rough_reflect(F3 perfect, float roughness, Xorwow& rng) {
    const float u1 = xw_uniform(rng);
    const float r1 = (float)((double)(u1 * 2.0f) * 3.141592653589793);
    const float r2 = xw_uniform(rng);
    const float z = 1.0f - 2.0f * r2;
    const float r = sqrtf(1.0f - z * z);
    const F3 sph = f3(r * __cosf(r1), r * __sinf(r1), z);
    return v_norm(v_add(perfect, v_scale(sph, roughness)));
}

// Returns true when the path continues with io.ray, false when it ends at this vertex.
// `id` >= 0 is the hit object, `t` the hit distance.
TRT_DEV bool shade_vertex(const SceneDev& sc, const RenderConsts& rc, PathVertexIO& io, int id, float t) {
    io.shadow = false;
    const float4* op = sc.objects + (size_t)id * 7;
    const F3 v0 = ld3(__ldg(op)), v1 = ld3(__ldg(op + 1)), v2 = ld3(__ldg(op + 2));
    F3 albedo = ld3(__ldg(op + 3));
    const F3 emission = ld3(__ldg(op + 4));
    const float4 mat = __ldg(op + 5);
    const int tex_id = f2i(__ldg(op + 6).x);
    const float metallic = mat.x, roughness = mat.y, ior = mat.z, transmission = mat.w;

    const F3 r_d = io.ray.d;
    const F3 x_hit = v_add(io.ray.o, v_scale(r_d, t));
    const F3 n = v_norm(v_cross(v_sub(v1, v0), v_sub(v2, v0)));
    const F3 nl = v_dot(n, r_d) < 0.f ? n : v_scale(n, -1.f);

    if (tex_id >= 0) {  // planar mapping, scale 0.01, V flipped (:465-481)
        float u, v;
        if (fabsf(n.y) > 0.9f)      { u = x_hit.x; v = x_hit.z; }
        else if (fabsf(n.x) > 0.9f) { u = x_hit.z; v = x_hit.y; }
        else                        { u = x_hit.x; v = x_hit.y; }
        u *= 0.01f;
        v *= 0.01f;
        v = 1.0f - v;
        // (a chain of selects instead of sc.tex[tex_id]: a dynamic index would force the whole parameter
        // struct into local memory)
        cudaTextureObject_t to = sc.tex[0];
        if (tex_id == 1) to = sc.tex[1];
        if (tex_id == 2) to = sc.tex[2];
        if (tex_id == 3) to = sc.tex[3];
        if (tex_id == 4) to = sc.tex[4];
        const float4 tx = tex2D<float4>(to, u, v);
        albedo = v_mul(albedo, f3(tx.x, tx.y, tx.z));
    }

    if (io.prev_mode != MODE_DIFF) io.rad = v_add(io.rad, v_mul(io.thr, emission));  // :489-495
    if (emission.x > 0.001f || emission.y > 0.001f || emission.z > 0.001f) return false;  // :497-499

    // lobe weights (:509-556)
    const float diffuse_suppression = powf(1.0f - metallic, 2.0f);
    float spec_attenuation = 1.0f - (roughness * roughness);
    if (spec_attenuation < 0.f) spec_attenuation = 0.f;
    const F3 F0 = v_add(v_scale(f3(0.04f, 0.04f, 0.04f), 1.0f - metallic), v_scale(albedo, metallic));
    const float cos_theta = fmaxf(v_dot(nl, v_scale(r_d, -1.0f)), 0.0f);
    const F3 F = v_add(F0, v_scale(v_sub(f3(1.f, 1.f, 1.f), F0), powf(1.0f - cos_theta, 5.0f)));
    const float F_avg = (F.x + F.y + F.z) / 3.0f;
    const float w_spec = F_avg * spec_attenuation;
    const float w_trans = (1.0f - F_avg) * transmission;
    const float albedo_max = fmaxf(albedo.x, fmaxf(albedo.y, albedo.z));
    float w_diff = (1.0f - F_avg) * (1.0f - transmission) * diffuse_suppression * albedo_max;
    float sum = w_spec + w_trans + w_diff;
    if (sum < 1e-5f) { w_diff = 1.0f; sum = 1.0f; }
    const float p_spec = w_spec / sum;
    const float p_trans = w_trans / sum;

    {  // Russian roulette (:559-565).  Written without a branch around the draw: an early return inside a
       // conditional makes the function exit the reconvergence point, and the lanes that played and the lanes
       // that did not would run the whole rest of the vertex one group after the other.
        const bool play = io.depth > rc.rr_threshold;
        float p = albedo_max;
        if (p < 0.05f) p = 0.05f;
        Xorwow drawn = io.rng;
        const float u = xw_uniform(drawn);
        if (play) io.rng = drawn;
        if (play && !(u < p)) return false;
        const float boost = play ? 1.0f / p : 1.0f;
        if (play) io.thr = v_scale(io.thr, boost);
    }

    const float rnd = xw_uniform(io.rng);

    if (rnd < p_spec) {  // specular (:571-589)
        const F3 perfect = v_sub(r_d, v_scale(v_scale(n, 2.f), v_dot(n, r_d)));
        const F3 nd = rough_reflect(perfect, roughness, io.rng);
        if (v_dot(nd, nl) <= 0.0f) return false;
        io.thr = v_scale(v_mul(io.thr, F), 1.0f / p_spec);
        io.ray.o = v_add(x_hit, v_scale(nl, 1e-3f));
        io.ray.d = nd;
        io.prev_mode = MODE_SPEC;
    } else if (rnd < p_spec + p_trans) {  // transmission (:592-648)
        const bool into = v_dot(n, nl) > 0.f;
        const float nnt = into ? 1.0f / ior : ior / 1.0f;
        const float ddn = v_dot(r_d, nl);
        const float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
        F3 nd;
        if (cos2t < 0.0f) {  // total internal reflection
            const F3 perfect = v_sub(r_d, v_scale(v_scale(n, 2.0f), v_dot(n, r_d)));
            nd = rough_reflect(perfect, roughness, io.rng);
        } else {
            nd = v_norm(v_sub(v_scale(r_d, nnt), v_scale(n, (into ? 1.0f : -1.0f) * (ddn * nnt + sqrtf(cos2t)))));
            if (roughness > 0.0f) {
                const float u1 = xw_uniform(io.rng);
                const float r1 = (float)((double)(u1 * 2.0f) * 3.141592653589793);
                const float r2 = xw_uniform(io.rng);
                const float z = 1.0f - 2.0f * r2;
                const float r = sqrtf(1.0f - z * z);
                nd = v_norm(v_add(nd, v_scale(f3(r * __cosf(r1), r * __sinf(r1), z), roughness)));
            }
        }
        io.ray.d = nd;
        io.ray.o = v_add(x_hit, v_scale(nd, 1e-4f));
        const float p_branch = (1.0f - p_spec) * transmission;
        if (p_branch > 1e-4f) io.thr = v_scale(v_mul(io.thr, albedo), 1.0f / p_branch);
        io.prev_mode = MODE_REFR;
    } else {  // diffuse with next-event estimation (:651-733)
        if (sc.n_lights > 0) {
            const int l_idx = (int)(xw_uniform(io.rng) * ((float)sc.n_lights - 0.001f));
            const float4* lp = sc.objects + (size_t)__ldg(sc.lights + l_idx) * 7;
            const F3 l0 = ld3(__ldg(lp)), l1 = ld3(__ldg(lp + 1)), l2 = ld3(__ldg(lp + 2));
            const F3 le = ld3(__ldg(lp + 4));
            const float r1 = xw_uniform(io.rng);
            const float r2 = xw_uniform(io.rng);
            const float sqr1 = sqrtf(r1);
            const float u = 1.0f - sqr1;
            const float v = sqr1 * (1.0f - r2);
            const F3 light_pos = v_add(v_add(v_scale(l0, u), v_scale(l1, v)), v_scale(l2, 1.0f - u - v));
            const F3 to_light = v_sub(light_pos, x_hit);
            float dist_sq = v_dot(to_light, to_light);
            if (dist_sq < 5.f) dist_sq = 5.f;
            const float dist = sqrtf(dist_sq);
            const F3 L = v_scale(to_light, 1.0f / dist);
            const float cos_s = v_dot(nl, L);
            const F3 lcross = v_cross(v_sub(l1, l0), v_sub(l2, l0));
            const F3 light_n = v_norm(lcross);
            const float cos_l = -v_dot(light_n, L);
            if (cos_s > 0.0f && cos_l > 0.0f) {
                const float area = sqrtf(lcross.x * lcross.x + lcross.y * lcross.y + lcross.z * lcross.z) * 0.5f;
                const float pdf = 1.0f / (area * (float)sc.n_lights);
                const float G = (cos_s * cos_l) / dist_sq;
                const F3 brdf = v_scale(albedo, (float)(1.0 / 3.141592653589793));
                const F3 contribution = v_scale(v_mul(le, brdf), G / pdf);
                io.shadow = true;
                io.shadow_ray.o = v_add(x_hit, v_scale(nl, 1e-3f));
                io.shadow_ray.d = L;
                io.shadow_max_dist = dist - 1e-2f;
                io.shadow_contrib = v_mul(io.thr, contribution);
            }
        }
        const F3 diffuse = v_scale(albedo, 1.0f - metallic);
        const float r1 = two_pi_times(xw_uniform(io.rng));
        const float r2 = xw_uniform(io.rng);
        const float r2s = sqrtf(r2);
        const F3 w = nl;
        const F3 up = fabsf(w.x) > 0.1f ? f3(0.f, 1.f, 0.f) : f3(1.f, 0.f, 0.f);
        const F3 u = v_norm(v_cross(up, w));
        const F3 v = v_cross(w, u);
        io.ray.d = v_norm(v_add(v_add(v_scale(v_scale(u, __cosf(r1)), r2s), v_scale(v_scale(v, __sinf(r1)), r2s)),
                                v_scale(w, sqrtf(1.f - r2))));
        const float p_diff = 1.0f - p_spec - (1.0f - p_spec) * transmission;
        io.thr = v_scale(v_mul(io.thr, diffuse), 1.0f / p_diff);
        io.ray.o = v_add(x_hit, v_scale(nl, 1e-3f));
        io.prev_mode = MODE_DIFF;
    }
    return true;
}

// Sample filter of reference :739-756.  Returns false when the sample is dropped.
TRT_DEV bool filter_sample(F3& rad) {
    if (isnan(rad.x) || isnan(rad.y) || isnan(rad.z) || isinf(rad.x) || isinf(rad.y) || isinf(rad.z)) return false;
    if (rad.x < 0.f) rad.x = 0.f;
    if (rad.y < 0.f) rad.y = 0.f;
    if (rad.z < 0.f) rad.z = 0.f;
    const float lum = (float)((double)rad.x * 0.21 + (double)rad.y * 0.71 + (double)rad.z * 0.07);
    if (lum > 100.f) rad = v_scale(rad, 100.f / lum);
    return true;
}

}  // namespace trt
