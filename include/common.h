// common.h -- float3 vector type shared by host and device code.
//
// Drop-in for the reference's include/common.h (struct Vec :24-97, make_vec :105,
// clamp :114, toInt :126).  Layout contract: 16 bytes,
// 16-byte aligned, members x,y,z at 0/4/8, 4 trailing pad bytes.
#pragma once

#include <cmath>
#include <cuda_runtime.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846f
#endif

#define TRT_HD __host__ __device__

struct __align__(16) Vec {
    float x, y, z;

    TRT_HD Vec operator+(const Vec& o) const { return Vec{x + o.x, y + o.y, z + o.z}; }
    TRT_HD Vec operator-(const Vec& o) const { return Vec{x - o.x, y - o.y, z - o.z}; }
    TRT_HD Vec operator*(float s) const { return Vec{x * s, y * s, z * s}; }

    // per-channel product (colour filtering)
    TRT_HD Vec mult(const Vec& o) const { return Vec{x * o.x, y * o.y, z * o.z}; }

    TRT_HD float dot(const Vec& o) const { return x * o.x + y * o.y + z * o.z; }

    TRT_HD Vec cross(const Vec& o) const {
        return Vec{y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x};
    }

    TRT_HD float norm_len() const { return sqrtf(x * x + y * y + z * z); }

    // in-place normalisation; a zero vector is left untouched
    TRT_HD Vec& norm() {
        const float l = sqrtf(x * x + y * y + z * z);
        if (l > 0) {
            const float r = 1.0f / l;
            x *= r;
            y *= r;
            z *= r;
        }
        return *this;
    }
};

static_assert(sizeof(Vec) == 16 && alignof(Vec) == 16, "Vec must be a padded float3");

TRT_HD inline Vec make_vec(float x, float y, float z) { return Vec{x, y, z}; }

// display helpers (host): saturate, then gamma 2.2 and 8-bit quantisation
inline float clamp(float v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }
inline int toInt(float v) { return int(pow(clamp(v), 1 / 2.2) * 255 + .5); }
