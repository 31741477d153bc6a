// aabb.h -- axis-aligned box with the slab test used by the reference-order
// traversal.  Drop-in for the reference's include/aabb.h (:5-6 wrappers,
// :11-46 empty/grow, :49-69 hit).  Layout: 32 bytes, min at 0, max at 16.
#pragma once
#include "common.h"

// ternary min/max: with a NaN operand the SECOND argument is returned, which the
// slab test below relies on (reference aabb.h:5-6).
TRT_HD inline float fmin_wrapper(float a, float b) { return a < b ? a : b; }
TRT_HD inline float fmax_wrapper(float a, float b) { return a > b ? a : b; }

struct __align__(16) AABB {
    Vec min;
    Vec max;

    TRT_HD static AABB empty() {
        AABB b;
        b.min = Vec{1e30f, 1e30f, 1e30f};
        b.max = Vec{-1e30f, -1e30f, -1e30f};
        return b;
    }

    TRT_HD void grow(Vec p) {
        min = Vec{fmin_wrapper(min.x, p.x), fmin_wrapper(min.y, p.y), fmin_wrapper(min.z, p.z)};
        max = Vec{fmax_wrapper(max.x, p.x), fmax_wrapper(max.y, p.y), fmax_wrapper(max.z, p.z)};
    }

    TRT_HD void grow(const AABB& b) {
        min = Vec{fmin_wrapper(min.x, b.min.x), fmin_wrapper(min.y, b.min.y), fmin_wrapper(min.z, b.min.z)};
        max = Vec{fmax_wrapper(max.x, b.max.x), fmax_wrapper(max.y, b.max.y), fmax_wrapper(max.z, b.max.z)};
    }

    // Slab test against the open interval (t_min, t_max); r_inv_d is 1/direction.
    // Plane distances are (plane - origin) * inv: a subtract followed by a
    // multiply, never an FMA -- the parity kernels pin exactly this sequence.
    __device__ bool hit(const Vec& r_o, const Vec& r_inv_d, float t_min, float t_max) const {
        float a = (min.x - r_o.x) * r_inv_d.x, b = (max.x - r_o.x) * r_inv_d.x;
        float lo = fmin_wrapper(a, b), hi = fmax_wrapper(a, b);
        a = (min.y - r_o.y) * r_inv_d.y;
        b = (max.y - r_o.y) * r_inv_d.y;
        lo = fmax_wrapper(lo, fmin_wrapper(a, b));
        hi = fmin_wrapper(hi, fmax_wrapper(a, b));
        a = (min.z - r_o.z) * r_inv_d.z;
        b = (max.z - r_o.z) * r_inv_d.z;
        lo = fmax_wrapper(lo, fmin_wrapper(a, b));
        hi = fmin_wrapper(hi, fmax_wrapper(a, b));
        return hi >= lo && hi > t_min && lo < t_max;
    }
};

static_assert(sizeof(AABB) == 32, "AABB is two padded float3");
