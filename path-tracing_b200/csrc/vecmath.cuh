// vecmath.cuh — small float3 / float2 algebra for the device code (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PT_DEV __device__ __forceinline__
#define PT_HD __host__ __device__ __forceinline__

namespace pt
{

struct vec2
{
    float x, y;
};
struct vec3
{
    float x, y, z;
};

PT_HD vec2 V2(float x, float y) { return vec2 { x, y }; }
PT_HD vec3 V3(float x, float y, float z) { return vec3 { x, y, z }; }
PT_HD vec3 V3(float s) { return vec3 { s, s, s }; }
PT_HD vec3 V3(float4 v) { return vec3 { v.x, v.y, v.z }; }

PT_HD vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
PT_HD vec2 operator-(vec2 a, vec2 b) { return { a.x - b.x, a.y - b.y }; }
PT_HD vec2 operator*(vec2 a, float s) { return { a.x * s, a.y * s }; }
PT_HD vec2 operator*(float s, vec2 a) { return { s * a.x, s * a.y }; }

PT_HD vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
PT_HD vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
PT_HD vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }
PT_HD vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
PT_HD vec3 operator/(vec3 a, vec3 b) { return { a.x / b.x, a.y / b.y, a.z / b.z }; }
PT_HD vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
PT_HD vec3 operator*(float s, vec3 a) { return { s * a.x, s * a.y, s * a.z }; }
PT_HD vec3 operator/(vec3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
PT_HD vec3 operator+(vec3 a, float s) { return { a.x + s, a.y + s, a.z + s }; }
PT_HD vec3 operator-(vec3 a, float s) { return { a.x - s, a.y - s, a.z - s }; }
PT_HD vec3 &operator+=(vec3 &a, vec3 b)
{
    a = a + b;
    return a;
}
PT_HD vec3 &operator-=(vec3 &a, vec3 b)
{
    a = a - b;
    return a;
}
PT_HD vec3 &operator*=(vec3 &a, vec3 b)
{
    a = a * b;
    return a;
}

PT_HD float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
PT_HD float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD vec3 cross(vec3 a, vec3 b)
{
    return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y };
}
PT_DEV float length(vec3 a) { return sqrtf(dot(a, a)); }
// GLSL.std.450 Normalize as glm spells it: v * inversesqrt(dot(v, v)); normalize(0) = NaN
PT_DEV vec3 normalize(vec3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
PT_DEV vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
PT_DEV float maxComponent(vec3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }
PT_DEV vec3 vmax(vec3 a, float b) { return { fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b) }; }
PT_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
PT_DEV vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }
// GLSL refract: zero vector on total internal reflection
PT_DEV vec3 refract(vec3 I, vec3 N, float eta)
{
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f)
        return V3(0.0f);
    return eta * I - (eta * d + sqrtf(k)) * N;
}
PT_DEV bool bad(float x) { return isnan(x) || isinf(x); }

// column-major 3x3 like GLSL mat3
struct mat3
{
    vec3 c0, c1, c2;
};
PT_HD vec3 mul(const mat3 &m, vec3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
// transpose(m) * v
PT_HD vec3 mulT(const mat3 &m, vec3 v) { return { dot(m.c0, v), dot(m.c1, v), dot(m.c2, v) }; }

} // namespace pt
