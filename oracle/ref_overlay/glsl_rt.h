/*
 * glsl_rt.h — GLSL language environment for the reference's shaders compiled as C++.
 *
 * TEST INFRASTRUCTURE (oracle/_ref/libglsl_ref.so, "the reference compiled here").  First of three
 * runtime headers the generated translation unit includes (see glsl2cpp.py):
 *   glsl_rt.h        language level: glm (the reference's own vendored copy) as the vector library,
 *                    plus the handful of implicit conversions GLSL has and glm lacks;
 *   glsl_rt_state.h  pipeline level: the descriptor-bound names, payloads and gl_* built-ins the stages
 *                    read, texture / image / traceRayEXT entry points;
 *   glsl_rt_api.h    the C ABI of libglsl_ref.so (stage drivers + per-function probes).
 *
 * Everything the shaders compute is compiled from the reference's text; this header adds no
 * arithmetic of its own except the documented conversions below.  Compile with -ffp-contract=off:
 * GLSL only fuses where the source says fma().
 */
#pragma once

#define GLM_FORCE_SWIZZLE
#define GLM_FORCE_SILENT_WARNINGS
#include <glm/glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstring>

namespace glslref
{

using namespace glm;

/* GLSL converts uvec2 -> vec2 implicitly (ray.glsl:19 `pixel + u`, :26 `pixelCenter / resolution`) */
inline vec2 operator+(uvec2 a, vec2 b) { return vec2(a) + b; }
inline vec2 operator/(vec2 a, uvec2 b) { return a / vec2(b); }
/* GLSL converts an int operand to float (tracing.glsl:106 `2 * (...)`, closestHit.rchit:81 `n *= -1`) */
template <length_t L> inline vec<L, float, defaultp> operator*(int s, vec<L, float, defaultp> const &v) { return (float)s * v; }
template <length_t L> inline vec<L, float, defaultp> operator*(vec<L, float, defaultp> const &v, int s) { return v * (float)s; }
template <length_t L> inline vec<L, float, defaultp> operator/(vec<L, float, defaultp> const &v, int s) { return v / (float)s; }
/* integer dot product (common.glsl:145 `dot(pixel, uvec2(1, resolution.x))`); glm's dot
 * is floating-point only */
inline uint dot(uvec2 a, uvec2 b) { return a.x * b.x + a.y * b.y; }
/* GLSL converts the int exponent to float (shading.glsl:52 `pow(x, 5)`, :104 `pow(eta, 2)`); a C++
 * pow(float, int) would promote to double.  Declaring pow here hides glm::pow for scalar calls, so the
 * float/float form is restated as the same std::pow glm::pow forwards to. */
inline float pow(float x, float y) { return std::pow(x, y); }
inline float pow(float x, int y) { return std::pow(x, (float)y); }

/* opaque handle types */
struct sampler2D
{
    uint32_t slot;
};
struct samplerCube
{
    uint32_t unused;
};
struct accelerationStructureEXT
{
    uint32_t unused;
};
struct image2D
{
    uint32_t unused;
};

} // namespace glslref
