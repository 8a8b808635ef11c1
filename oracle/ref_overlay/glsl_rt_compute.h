/*
 * glsl_rt_compute.h — the pipeline state the reference's COMPUTE stages see (postprocess.comp,
 * bloomDownsample.comp, bloomUpsample.comp, composition.comp, toneMapping.comp, skinning.comp), the
 * compute counterpart of glsl_rt_state.h.  TEST INFRASTRUCTURE (oracle/_ref/libglsl_comp_ref.so).
 *
 * What the shaders compute is compiled from the reference's text.  What the Vulkan implementation does around
 * them is the harness's and stays PARITY UNPINNED, exactly as for the ray-tracing stages:
 *   - storage images: a store into an RGBA16F image rounds to binary16 (round to nearest even; done here with
 *     the compiler's _Float16, i.e. independently of the oracle's bit manipulation), RGBA32F stores keep the float;
 *   - the bloom sampler (Renderer.cpp:114-119): linear filter, clamp to edge, one level — unnormalised
 *     coordinate u * size - 0.5, weights in fp32.
 */
#pragma once

#include "../../include/pt_core.h"

#include <algorithm>
#include <vector>

namespace glslref
{

/* GLSL converts the uint operand to float (postprocess.comp:22 `accColor / mainUniform.TotalSamples`) */
template <length_t L> inline vec<L, float, defaultp> operator/(vec<L, float, defaultp> const &v, uint s) { return v / (float)s; }
inline uint nonuniformEXT(uint i) { return i; }

/* one storage image / sampled image: RGBA floats; half = the format is RGBA16F */
struct ImageData
{
    int w = 0, h = 0;
    bool half = true;
    std::vector<vec4> px;
    void resize(int w_, int h_, bool half_)
    {
        w = w_, h = h_, half = half_;
        px.assign((size_t)w * h, vec4(0.0f));
    }
};
static thread_local ImageData *tls_images[64];

inline float roundToHalf(float f) { return (float)(_Float16)f; }

inline vec4 imageLoad(image2D img, ivec2 p) { return tls_images[img.unused]->px[(size_t)p.y * tls_images[img.unused]->w + p.x]; }
inline void imageStore(image2D img, ivec2 p, vec4 v)
{
    ImageData &d = *tls_images[img.unused];
    if (p.x < 0 || p.y < 0 || p.x >= d.w || p.y >= d.h)
        return; /* out-of-bounds image stores are discarded (the dispatch is rounded up to 32 x 32 groups) */
    if (d.half)
        v = vec4(roundToHalf(v.x), roundToHalf(v.y), roundToHalf(v.z), roundToHalf(v.w));
    d.px[(size_t)p.y * d.w + p.x] = v;
}
inline ivec2 imageSize(image2D img) { return ivec2(tls_images[img.unused]->w, tls_images[img.unused]->h); }
inline ivec2 textureSize(sampler2D s, int) { return ivec2(tls_images[s.slot]->w, tls_images[s.slot]->h); }
inline vec4 texture(sampler2D s, vec2 uv)
{
    const ImageData &d = *tls_images[s.slot];
    const float x = uv.x * (float)d.w - 0.5f, y = uv.y * (float)d.h - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int x0 = std::clamp((int)fx0, 0, d.w - 1), x1 = std::clamp((int)fx0 + 1, 0, d.w - 1);
    const int y0 = std::clamp((int)fy0, 0, d.h - 1), y1 = std::clamp((int)fy0 + 1, 0, d.h - 1);
    const vec4 top = d.px[(size_t)y0 * d.w + x0] * (1.0f - fx) + d.px[(size_t)y0 * d.w + x1] * fx;
    const vec4 bot = d.px[(size_t)y1 * d.w + x0] * (1.0f - fx) + d.px[(size_t)y1 * d.w + x1] * fx;
    return top * (1.0f - fy) + bot * fy;
}

static thread_local uvec3 gl_GlobalInvocationID;

/* `uint[] inIndices` of skinning.comp:17-19 */
struct UintBuffer
{
    const uint *v = nullptr;
    uint n = 0;
    uint length() const { return n; }
    uint operator[](uint i) const { return v[i]; }
};

/* ---- descriptor-bound names per stage (the declarations rule 2 drops) ---- */
#define GLSL_COMPUTE_STATE_comp_post                                                                                   \
    static thread_local image2D u_AccumulationImage, u_PostProcessImage, u_BloomImage;                                 \
    static thread_local PostProcessingUniformData mainUniform;
#define GLSL_COMPUTE_STATE_comp_down                                                                                   \
    static thread_local sampler2D u_BloomSampler[MaxBloomMipmapLevel + 1];                                             \
    static thread_local image2D u_BloomMipmaps[MaxBloomMipmapLevel + 1];                                               \
    static thread_local struct                                                                                         \
    {                                                                                                                  \
        uint mipmapLevel;                                                                                              \
    } pushConstants;
#define GLSL_COMPUTE_STATE_comp_up GLSL_COMPUTE_STATE_comp_down
#define GLSL_COMPUTE_STATE_comp_compose                                                                                \
    static thread_local image2D u_PostProcessImage, u_BloomImage;                                                      \
    static thread_local PostProcessingUniformData mainUniform;
#define GLSL_COMPUTE_STATE_comp_tone                                                                                   \
    static thread_local uint s_ToneMappingMode;                                                                        \
    static thread_local image2D u_Image;
#define GLSL_COMPUTE_STATE_comp_skin                                                                                   \
    static thread_local SkinningPushConstants pc;                                                                      \
    static thread_local const mat3x4 *boneTransforms;                                                                  \
    static thread_local UintBuffer inIndices;

} // namespace glslref
