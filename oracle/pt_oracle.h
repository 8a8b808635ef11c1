/*
 * pt_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (path-tracing_b200/, include/pt_core.h) never does.
 *
 * The oracle is a literal, scalar CPU restatement of the reference's GLSL hot
 * path (PT/Shaders/{common,ray,tracing,shading,bsdf,sampling,material}.glsl,
 * raygen.rgen, closestHit.rchit, anyhit.rahit, occlusionAnyhit.rahit,
 * miss.rmiss, occlusion.rmiss), one thread of control per pixel like the
 * reference's raygen invocation.  See pt_oracle.cpp for per-function
 * citations and for what is and is not pinned by the reference's own tests.
 *
 * It consumes the same scene / parameter PODs as the product ABI (pt_core.h).
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include "../include/pt_core.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pto_scene pto_scene;

/* Extra per-run counters the roofline formula of SURVEY §8(d) needs. */
typedef struct pto_counters {
    uint64_t rays_closest;
    uint64_t rays_shadow;
    uint64_t samples;
    uint64_t hits;
    uint64_t box_tests_closest;
    uint64_t tri_tests_closest;
    uint64_t box_tests_shadow;
    uint64_t tri_tests_shadow;
    uint64_t alpha_tests_closest;
    uint64_t alpha_tests_shadow;
    uint64_t texel_fetches; /* texels read by the material textureGrad fetches */
    uint64_t restarts;
} pto_counters;

PT_API pto_scene *pto_scene_create(const pt_scene_desc *scene);
/* The same with TextureUploader's limits (TextureUploader.cpp:408-415, 551-569; like the core's tuning keys
 * "max_texture_size" and "texture_budget_mb"): RGBA8 scene textures beyond the maximum extent are scaled down by an integer
 * factor with a linear blit before their mips are generated.  pto_scene_create = (4096, 0 = ForceFullTextureSize). */
PT_API pto_scene *pto_scene_create_limits(const pt_scene_desc *scene, uint32_t max_texture_size, uint64_t texture_budget_bytes);
PT_API void pto_scene_destroy(pto_scene *scene);
PT_API uint64_t pto_scene_triangle_count(const pto_scene *scene);
/* Sampler state like pt_set_sampler: maximum anisotropy of textureGrad, 1 (isotropic trilinear, the default) .. 16 (what
 * the reference's sampler asks for: anisotropy at the device maximum, Renderer.cpp:103-112). */
PT_API int32_t pto_scene_set_sampler(pto_scene *scene, uint32_t max_anisotropy);

/* accum: width*height*4 floats, updated in place exactly like imageLoad/imageStore in
 * raygen.rgen:115-117 (rgb += radiance, a = 1), for frames TotalSamples = first_sample ..
 * first_sample + sample_count - 1 with SampleCount = 1.  Only pixels inside the tiles are
 * touched (tiles == NULL: all).  threads <= 0 uses std::thread::hardware_concurrency(). */
PT_API int32_t pto_render(const pto_scene *scene, const pt_render_params *params, uint32_t width,
                          uint32_t height, uint32_t first_sample, uint32_t sample_count, const pt_tile *tiles,
                          uint32_t tile_count, float *accum, int32_t threads, pto_counters *out_counters);

/* The same with the reference's Release-profile frame structure (SamplesPerFrame > 1, Renderer.cpp:1688-1700):
 * frame f runs raygen.rgen:36-118 once per pixel with SampleCount = samples_per_frame and
 * TotalSamples = first_sample + f * samples_per_frame — ONE RNG stream and ONE radiance sum per frame. */
PT_API int32_t pto_render_frames(const pto_scene *scene, const pt_render_params *params, uint32_t width, uint32_t height,
                                 uint32_t first_sample, uint32_t frame_count, uint32_t samples_per_frame,
                                 const pt_tile *tiles, uint32_t tile_count, float *accum, int32_t threads,
                                 pto_counters *out_counters);

/* Hooks for oracle/_ref/libglsl_ref.so (the reference's shaders compiled as C++, oracle/ref_overlay/build_glsl.sh):
 * the two services the reference leaves to the Vulkan implementation.  pto_trace_anyhit is the oracle's own
 * traversal with the any-hit stage handed to the caller (1 = accept, 0 = ignoreIntersectionEXT; NULL = the
 * oracle's inline any-hit); returns 1 and fills *out on a hit.  pto_sky_sample: kind 0 = texture(skybox2D, uv),
 * kind 1 = texture(skyboxCube, dir). */
typedef int32_t (*pto_anyhit_fn)(void *ctx, uint32_t instance, uint32_t geometry, uint32_t primitive, float t, float b1,
                                 float b2);
PT_API int32_t pto_trace_anyhit(const pto_scene *scene, const float *org, const float *dir, float tmin, float tmax,
                                uint32_t terminate_on_first_hit, pto_anyhit_fn anyhit, void *ctx, pt_hit *out);
/* the same with ray flags: bit 0 = gl_RayFlagsOpaqueEXT (no candidate reaches the any-hit stage), bit 1 =
 * gl_RayFlagsCullBackFacingTrianglesEXT — the debug pipeline's primary rays (Debug/debugRaygen.rgen:28-35) */
PT_API int32_t pto_trace_anyhit_flags(const pto_scene *scene, const float *org, const float *dir, float tmin, float tmax,
                                      uint32_t terminate_on_first_hit, uint32_t flags, pto_anyhit_fn anyhit, void *ctx, pt_hit *out);
PT_API int32_t pto_sky_sample(const pto_scene *scene, uint32_t kind, const float *in3, float *out4);
/* closestHit.rchit:52-161 on `count` given hits: rays6 = world ray origin.xyz, direction.xyz; payloads are the 36
 * words of Shaders::Payload (ShaderRendererTypes.incl:101-118; RngState as bits). */
PT_API int32_t pto_closest_hit(const pto_scene *scene, const pt_render_params *params, uint32_t count, const pt_hit *hits,
                               const float *rays6, const float *payload_in, float *payload_out);

PT_API int32_t pto_first_hit_aov(const pto_scene *scene, const pt_render_params *params, uint32_t width,
                                 uint32_t height, pt_hit *out_hits);
PT_API int32_t pto_trace_closest(const pto_scene *scene, const pt_ray *rays, uint64_t ray_count,
                                 pt_hit *out_hits);
PT_API int32_t pto_trace_occlusion(const pto_scene *scene, const pt_ray *rays, uint64_t ray_count,
                                   uint8_t *out_occluded);

/* Brute-force double-precision closest hit over all triangles (no BVH, Moller-Trumbore in
 * fp64) — the independent check of the oracle's own BVH + fp32 watertight test. */
PT_API int32_t pto_trace_closest_bruteforce_f64(const pto_scene *scene, const pt_ray *rays,
                                                uint64_t ray_count, pt_hit *out_hits, double *out_t);

/* Same modes / record layouts as pt_test_shading in pt_core.h. */
PT_API int32_t pto_test_shading(uint32_t mode, const float *input, float *output, uint32_t count);

/* Texture sampler probes (slot indexes the bindless array: 0-8 built in, 9+ scene). */
PT_API int32_t pto_texture_info(const pto_scene *scene, uint32_t slot, uint32_t *out_width,
                                uint32_t *out_height, uint32_t *out_levels);
/* Copies RGBA8 level `level` of an 8-bit texture (width*height*4 bytes). */
PT_API int32_t pto_texture_level(const pto_scene *scene, uint32_t slot, uint32_t level, uint8_t *out_rgba8);
/* records: in uv.xy, ddx.xy, ddy.xy (6 floats) -> out rgba; use_grad = 0 samples LOD 0 (texture()). */
PT_API int32_t pto_texture_sample(const pto_scene *scene, uint32_t slot, const float *in6, float *out4,
                                  uint32_t count, int32_t use_grad);

/* Post-processing + output chain (pt_oracle_post.cpp; same parameters and formats as pt_postprocess)
 * applied to a host accumulation image of width*height float4 sums. */
PT_API int32_t pto_postprocess(const float *accum, uint32_t width, uint32_t height, const pt_postprocess_params *params,
                               uint32_t total_samples, uint32_t output_format, void *out_pixels);
/* The debug pipeline (same parameters as pt_debug_render): width*height RGBA floats. */
PT_API int32_t pto_debug_render(const pto_scene *scene, const pt_render_params *params, const pt_debug_params *debug,
                                uint32_t width, uint32_t height, float *out_rgba);
/* binary32 -> binary16 -> binary32 (round to nearest even) of n values */
PT_API int32_t pto_round_half(const float *in, float *out, uint64_t n);

/* skinning.comp:21-50 on `count` animated vertices (the probe tests/test_oracle_vs_glsl_compute.py compares with the
 * compiled shader): out[i] = the skinned vertex of animated[indices[i]] */
PT_API int32_t pto_skin_vertices(const pt_animated_vertex *animated, const uint32_t *indices, uint64_t count,
                                 const float *bone_transforms, uint32_t bone_count, pt_vertex *out);

#ifdef __cplusplus
}
#endif
#endif
