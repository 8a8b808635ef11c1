/*
 * glsl_rt_state.h — the pipeline state the reference's shader stages see (second runtime header,
 * included after the reference's type headers; see glsl_rt.h).  TEST INFRASTRUCTURE.
 *
 * One set of thread-local variables stands for the descriptor set, the ray payloads and the gl_*
 * built-ins of ONE shader invocation chain (raygen -> traceRayEXT -> any-hit / closest-hit / miss).
 * The two things the reference leaves to the Vulkan implementation — acceleration-structure traversal
 * and texture filtering — are reached through callbacks the test harness points at the CPU oracle
 * (pto_trace_anyhit, pto_texture_sample, pto_sky_sample): PARITY UNPINNED for those two, exactly as
 * DESIGN.md §4 says; everything else is the reference's own arithmetic.
 */
#pragma once

#include "../../include/pt_core.h"

extern "C" {
/* candidate hit handed to the any-hit stage; returns 1 = accept, 0 = ignoreIntersectionEXT */
typedef int32_t (*glr_anyhit_fn)(void *ctx, uint32_t instance, uint32_t geometry, uint32_t primitive, float t,
                                 float b1, float b2);
typedef struct glr_callbacks
{
    void *user; /* pto_scene* */
    /* closest hit (terminate_on_first_hit = 0) or any hit (1) in (tmin, tmax); non-opaque candidates go
     * through anyhit(ctx, ...).  Returns 1 and fills *out when something was hit. */
    int32_t (*trace)(const void *user, const float *org, const float *dir, float tmin, float tmax,
                     uint32_t terminate_on_first_hit, glr_anyhit_fn anyhit, void *ctx, pt_hit *out);
    /* (user, slot, {uv, ddx, ddy}, out rgba, count, use_grad) */
    int32_t (*texture)(const void *user, uint32_t slot, const float *in6, float *out4, uint32_t count, int32_t use_grad);
    /* kind 0: 2-D sky at uv (in3[0..1]); kind 1: cube sky along direction in3 */
    int32_t (*sky)(const void *user, uint32_t kind, const float *in3, float *out4);
    /* `trace` with ray flags (the debug pipeline): bit 0 = gl_RayFlagsOpaqueEXT, bit 1 = gl_RayFlagsCullBackFacingTrianglesEXT */
    int32_t (*trace_flags)(const void *user, const float *org, const float *dir, float tmin, float tmax,
                           uint32_t terminate_on_first_hit, uint32_t flags, glr_anyhit_fn anyhit, void *ctx, pt_hit *out);
} glr_callbacks;
}

namespace glslref
{

static_assert(sizeof(Vertex) == 56 && sizeof(MetallicRoughnessMaterial) == 96 && sizeof(SpecularGlossinessMaterial) == 96 &&
                  sizeof(PhongMaterial) == 96 && sizeof(PointLight) == 48 && sizeof(DirectionalLight) == 32 &&
                  sizeof(Payload) == 144 && sizeof(mat3x4) == 48 && sizeof(DebugPayload) == 48,
              "the GLSL branch of the dual headers must have the host layout (PTT/PaddingTest.cpp)");

struct SamplerArray
{
    sampler2D operator[](uint i) const { return sampler2D { i }; }
};

/* ---- descriptor set (raygen.rgen:8-15, closestHit.rchit:10-34, miss.rmiss:10-12) ---- */
static thread_local accelerationStructureEXT u_TopLevelAS;
static thread_local image2D u_Image;
static thread_local RaygenUniformData mainUniform;
static thread_local SamplerArray textures;
static thread_local const mat3x4 *transforms;
static thread_local const Geometry *geometries;
static thread_local const MetallicRoughnessMaterial *metallicRoughnessMaterials;
static thread_local const SpecularGlossinessMaterial *specularGlossinessMaterials;
static thread_local const PhongMaterial *phongMaterials;
static thread_local uint u_LightCount;
static thread_local DirectionalLight u_DirectionalLight;
static thread_local const PointLight *u_Lights;
static thread_local sampler2D skybox2D;
static thread_local samplerCube skyboxCube;
/* specialisation constants */
static thread_local uint s_HitFlags;
static thread_local uint s_MissFlags;
/* shader record, payloads, hit attributes */
static thread_local SBTBuffer sbt;
static thread_local Payload payload;
static thread_local bool isOccluded;
static thread_local vec3 attribs;
/* built-ins */
static thread_local uvec3 gl_LaunchIDEXT;
static thread_local uvec3 gl_LaunchSizeEXT;
static thread_local int gl_PrimitiveID;
static thread_local vec3 gl_WorldRayOriginEXT;
static thread_local vec3 gl_WorldRayDirectionEXT;
static thread_local float gl_RayTmaxEXT;
static thread_local mat3x4 gl_ObjectToWorld3x4EXT;
static thread_local int gl_GeometryIndexEXT;
static thread_local int gl_InstanceID;
const uint gl_RayFlagsNoneEXT = 0u;
const uint gl_RayFlagsOpaqueEXT = 1u;
const uint gl_RayFlagsTerminateOnFirstHitEXT = 4u;
const uint gl_RayFlagsCullBackFacingTrianglesEXT = 16u;
/* the debug pipeline (Debug/debugRaygen.rgen:8-18, debugClosestHit.rchit:9-10): its payload and specialisation constants */
static thread_local DebugPayload dbg_payload;
static thread_local uint s_RenderMode;
static thread_local uint s_RaygenFlags;
static thread_local uint s_HitGroupFlags;
static thread_local bool tls_debug_pipeline;

/* ---- harness side ---- */
struct Scene;
static thread_local const Scene *tls_scene;
static thread_local float *tls_image; /* rgba32f accumulation image, row-major */
static thread_local bool tls_ignore;
static thread_local uint64_t tls_trace_calls;

inline void glsl_ignore_intersection() { tls_ignore = true; }

vec4 texture(sampler2D s, vec2 uv);
vec4 textureGrad(sampler2D s, vec2 uv, vec2 dPdx, vec2 dPdy);
vec4 texture(samplerCube s, vec3 dir);
vec4 imageLoad(image2D img, ivec2 p);
void imageStore(image2D img, ivec2 p, vec4 v);
void traceRayEXT(accelerationStructureEXT tlas, uint rayFlags, uint cullMask, uint sbtRecordOffset, uint sbtRecordStride,
                 uint missIndex, vec3 origin, float tmin, vec3 direction, float tmax, int payloadLocation);

} // namespace glslref
