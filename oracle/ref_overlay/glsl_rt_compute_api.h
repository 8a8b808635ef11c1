/*
 * glsl_rt_compute_api.h — C ABI of oracle/_ref/libglsl_comp_ref.so (included after the generated compute stages;
 * see glsl_rt_compute.h).  TEST INFRASTRUCTURE: used only by tests/ to pin the CPU oracle's post-process chain
 * (oracle/pt_oracle_post.cpp) and skinning (oracle/pt_oracle.cpp skinVertex) to the reference's own shader text.
 *
 *   glc_postprocess   the dispatch sequence of Renderer::RecordPostProcessCommands (PT/Renderer/Renderer.cpp:928-1060)
 *                     + the in-place toneMapping.comp of RecordSaveOutputCommands (:1205-1250), every thread of every
 *                     dispatch running the stage's compiled main();
 *   glc_skin          skinning.comp main() per output vertex (Renderer::RecordSkinningCommands, :854-890).
 */
#pragma once

#include <cstring>

namespace glslref
{

static uint32_t mipLevelCount(uint32_t w, uint32_t h) /* PT/Renderer/Image.cpp:14-17 */
{
    uint32_t levels = 1;
    for (uint32_t m = std::max(w, h); m > 1; m >>= 1)
        levels++;
    return levels;
}

template <class Main> static void dispatch(int w, int h, Main main_)
{
    /* ceil(extent / 32) groups of 32 x 32 threads (Renderer.cpp:941-943): threads beyond the image run too */
    const int gw = (w + 31) / 32 * 32, gh = (h + 31) / 32 * 32;
    for (int y = 0; y < gh; y++)
        for (int x = 0; x < gw; x++)
        {
            gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
            main_();
        }
}

} // namespace glslref

extern "C" {

#define GLC_EXPORT __attribute__((visibility("default")))

/* accum: width x height RGBA32F.  out_composed / out_final: width x height RGBA floats = the RGBA16F values of the
 * post-process image after composition.comp and after toneMapping.comp (either may be NULL). */
GLC_EXPORT int32_t glc_postprocess(const float *accum, uint32_t width, uint32_t height, uint32_t total_samples, float exposure,
                                   float bloom_threshold, float bloom_intensity, uint32_t tone_mapping_mode, float *out_bloom0,
                                   float *out_composed, float *out_final)
{
    using namespace glslref;
    if (!accum || width == 0 || height == 0)
        return -1;
    const uint32_t levels = mipLevelCount(width, height);
    if (levels > MaxBloomMipmapLevel + 1)
        return -2;
    ImageData acc, pp;
    std::vector<ImageData> bloom(levels);
    acc.resize((int)width, (int)height, false);
    std::memcpy(acc.px.data(), accum, (size_t)width * height * 16);
    pp.resize((int)width, (int)height, true);
    for (uint32_t l = 0; l < levels; l++)
        bloom[l].resize((int)std::max(1u, width >> l), (int)std::max(1u, height >> l), true);
    /* image table: 0 = accumulation, 1 = post-process, 2 + l = bloom level l */
    tls_images[0] = &acc;
    tls_images[1] = &pp;
    for (uint32_t l = 0; l < levels; l++)
        tls_images[2 + l] = &bloom[l];
    const PostProcessingUniformData uniform = { total_samples, exposure, bloom_threshold, bloom_intensity };

    comp_post::u_AccumulationImage = image2D { 0 };
    comp_post::u_PostProcessImage = image2D { 1 };
    comp_post::u_BloomImage = image2D { 2 };
    comp_post::mainUniform = uniform;
    dispatch((int)width, (int)height, comp_post::main_);

    /* Renderer.cpp:955-1039; frames smaller than 8 pixels get no bloom passes (the reference's unsigned
     * `levels - 3` would index levels that do not exist) */
    const uint32_t maxMip = levels > 3 ? std::min(levels - 3, 12u) : 1u;
    for (uint32_t l = 0; l < levels; l++)
    {
        comp_down::u_BloomSampler[l] = sampler2D { 2 + l };
        comp_down::u_BloomMipmaps[l] = image2D { 2 + l };
        comp_up::u_BloomSampler[l] = sampler2D { 2 + l };
        comp_up::u_BloomMipmaps[l] = image2D { 2 + l };
    }
    for (uint32_t i = 0; i + 1 < maxMip; i++)
    {
        comp_down::pushConstants.mipmapLevel = i;
        dispatch(bloom[i + 1].w, bloom[i + 1].h, comp_down::main_);
    }
    for (uint32_t i = maxMip - 1; i > 0; i--)
    {
        comp_up::pushConstants.mipmapLevel = i;
        dispatch(bloom[i - 1].w, bloom[i - 1].h, comp_up::main_);
    }
    if (out_bloom0)
        std::memcpy(out_bloom0, bloom[0].px.data(), (size_t)width * height * 16);

    comp_compose::u_PostProcessImage = image2D { 1 };
    comp_compose::u_BloomImage = image2D { 2 };
    comp_compose::mainUniform = uniform;
    dispatch((int)width, (int)height, comp_compose::main_);
    if (out_composed)
        std::memcpy(out_composed, pp.px.data(), (size_t)width * height * 16);

    /* the blit into OutputSaver's RGBA16F linear image copies the halves; toneMapping.comp runs in place on it */
    comp_tone::s_ToneMappingMode = tone_mapping_mode;
    comp_tone::u_Image = image2D { 1 };
    dispatch((int)width, (int)height, comp_tone::main_);
    if (out_final)
        std::memcpy(out_final, pp.px.data(), (size_t)width * height * 16);
    return 0;
}

/* animated: animated_count x 22 floats (AnimatedVertex as the vec2 buffer getAnimatedVertex reads: bone indices as
 * bits); in_indices: out_count indices into it; bones: bone_count x 12 floats (mat3x4); out_vertices: out_count x 14. */
GLC_EXPORT int32_t glc_skin(const float *animated, uint32_t animated_count, const uint32_t *in_indices, uint32_t out_count,
                            const float *bones, uint32_t bone_count, float *out_vertices)
{
    using namespace glslref;
    if (!animated || !in_indices || !bones || !out_vertices || bone_count == 0)
        return -1;
    (void)animated_count;
    comp_skin::pc.inAnimatedVertices.v = (vec2 *)animated;
    comp_skin::pc.outVertices.v = (vec2 *)out_vertices;
    comp_skin::boneTransforms = (const mat3x4 *)bones;
    comp_skin::inIndices.v = in_indices;
    comp_skin::inIndices.n = out_count;
    /* ceil(count / 256) groups of 256 threads (Renderer.cpp:883-886) */
    const uint32_t threads = (out_count + 255u) / 256u * 256u;
    for (uint32_t i = 0; i < threads; i++)
    {
        gl_GlobalInvocationID = uvec3(i, 0u, 0u);
        comp_skin::main_();
    }
    return 0;
}

} // extern "C"
