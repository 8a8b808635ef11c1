#include "HeadlessRenderer.h"

#include <stb_image_write.h>

#include <cmath>
#include <cstring>
#include <format>

#include "Core/Core.h"


namespace PathTracing
{

HeadlessRenderer::HeadlessRenderer(int cudaDevice)
{
    const pt_status status = pt_context_create(cudaDevice, &m_Context);
    if (status != PT_OK)
        throw error(std::format("pt_context_create failed ({}): {}", status, pt_last_error(nullptr)));
}

HeadlessRenderer::~HeadlessRenderer()
{
    pt_context_destroy(m_Context);
}

void HeadlessRenderer::Check(pt_status status, const char *what)
{
    if (status != PT_OK)
        throw error(std::format("{} failed ({}): {}", what, status, pt_last_error(m_Context)));
}

void HeadlessRenderer::ResetAccumulation()
{
    m_TotalSamples = 0;
    if (m_Width != 0 && m_Height != 0)
        Check(pt_render_begin(m_Context, m_Width, m_Height), "pt_render_begin");
}

void HeadlessRenderer::UpdateSceneData(const std::shared_ptr<Scene> &scene, bool updated)
{
    if (updated)
        ResetAccumulation();

    if (m_Scene == scene)
    {
        /* Renderer::Render's per-frame work for animated scenes: the light uniform rewrite
         * (Renderer.cpp:1719-1726) and AccelerationStructure::RecordUpdateCommands (:1753-1754),
         * fed by what Scene::Update moved (Scene.cpp:52-83). */
        if (updated && scene->HasAnimations())
        {
            std::vector<float> transforms;
            for (const ModelInstance &instance : scene->GetModelInstances())
            {
                const float *m = reinterpret_cast<const float *>(&instance.Transform);
                transforms.insert(transforms.end(), m, m + 12);
            }
            const auto lights = scene->GetPointLights();
            const Shaders::DirectionalLight directional = scene->GetDirectionalLight();
            pt_scene_update_desc update = {};
            update.instance_transforms = transforms.data();
            update.instance_count = static_cast<uint32_t>(transforms.size() / 12);
            update.point_lights = reinterpret_cast<const pt_point_light *>(lights.data());
            update.point_light_count = static_cast<uint32_t>(lights.size());
            update.directional_light = reinterpret_cast<const pt_directional_light *>(&directional);
            /* Renderer::RecordSkinningCommands (Renderer.cpp:1750-1751) */
            const auto bones = scene->GetBoneTransforms();
            if (scene->HasSkeletalAnimations())
            {
                update.bone_transforms = reinterpret_cast<const float *>(bones.data());
                update.bone_count = static_cast<uint32_t>(bones.size());
            }
            Check(pt_scene_update(m_Context, &update), "pt_scene_update");
        }
        return;
    }

    m_Scene = scene;
    const auto flat = FlattenScene(*scene);
    Check(pt_scene_upload(m_Context, &flat->Desc), "pt_scene_upload");
    ResetAccumulation();
}

void HeadlessRenderer::OnResize(uint32_t width, uint32_t height)
{
    m_Width = width;
    m_Height = height;
    ResetAccumulation();
}

void HeadlessRenderer::SetSettings(const PathTracingSettings &settings)
{
    m_PathTracing = settings;
    ResetAccumulation();
}

void HeadlessRenderer::SetSettings(const PostProcessSettings &settings)
{
    m_PostProcess = settings;
}

pt_render_params HeadlessRenderer::MakeRenderParams()
{
    if (m_Scene == nullptr)
        throw error("HeadlessRenderer: no scene");

    /* Renderer.cpp:1686-1694 */
    Camera &camera = m_Scene->GetActiveCamera();
    camera.OnResize(m_Width, m_Height);
    const glm::mat4 invView = camera.GetInvViewMatrix();
    const glm::mat4 invProj = camera.GetInvProjectionMatrix();

    pt_render_params params = {};
    std::memcpy(params.view_inverse, &invView, sizeof(params.view_inverse));
    std::memcpy(params.proj_inverse, &invProj, sizeof(params.proj_inverse));
    params.bounce_count = m_PathTracing.BounceCount;
    params.lens_radius = m_PathTracing.LensRadius;
    params.focal_distance = m_PathTracing.FocalDistance;
    /* UpdateScenePipelineConfig: specialisation constants follow the scene (Renderer.cpp:711-754) */
    params.miss_flags = std::holds_alternative<Skybox2D>(m_Scene->GetSkybox())     ? PT_MISS_FLAGS_SKYBOX_2D
                        : std::holds_alternative<SkyboxCube>(m_Scene->GetSkybox()) ? PT_MISS_FLAGS_SKYBOX_CUBE
                                                                                   : PT_MISS_FLAGS_NONE;
    params.hit_flags = m_Scene->HasDxNormalTextures() ? PT_HIT_FLAGS_DX_NORMAL_TEXTURES : PT_HIT_FLAGS_NONE;

    return params;
}

void HeadlessRenderer::SetDebugRaytracingPipeline(const DebugRaytracingPipelineConfig &config)
{
    m_Debug = config;
}

std::vector<float> HeadlessRenderer::RenderDebug()
{
    pt_render_params params = MakeRenderParams();
    params.miss_flags = m_Debug.MissFlags;
    const pt_debug_params debug = { m_Debug.RenderMode, m_Debug.RaygenFlags, m_Debug.HitGroupFlags };
    std::vector<float> pixels(static_cast<size_t>(m_Width) * m_Height * 4);
    Check(pt_debug_render(m_Context, &params, &debug, m_Width, m_Height, pixels.data()), "pt_debug_render");
    return pixels;
}

void HeadlessRenderer::SetSamplesPerFrame(uint32_t samplesPerFrame)
{
    if (samplesPerFrame == 0)
        throw error("HeadlessRenderer: SamplesPerFrame must be at least 1");
    m_SamplesPerFrame = samplesPerFrame;
}

void HeadlessRenderer::Render(uint32_t frames)
{
    /* Renderer.cpp:1688-1700: TotalSamples = samples accumulated so far, SampleCount = SamplesPerFrame */
    const pt_render_params params = MakeRenderParams();
    Check(pt_render_frames(m_Context, &params, m_TotalSamples, frames, m_SamplesPerFrame, nullptr, 0), "pt_render_frames");
    m_TotalSamples += frames * m_SamplesPerFrame;
}

std::vector<float> HeadlessRenderer::ReadAccumulation()
{
    std::vector<float> pixels(static_cast<size_t>(m_Width) * m_Height * 4);
    Check(pt_readback(m_Context, pixels.data(), pixels.size() * sizeof(float)), "pt_readback");
    return pixels;
}

/* Renderer::RecordPostProcessCommands + RecordSaveOutputCommands + OutputSaver::WriteImage
 * (Renderer.cpp:928-1060, 1205-1250; OutputSaver.cpp:227-253): the whole chain runs in the core. */
std::vector<uint8_t> HeadlessRenderer::PostProcessSrgb8()
{
    const pt_postprocess_params params = { m_PostProcess.Exposure, m_PostProcess.BloomThreshold,
                                           m_PostProcess.BloomIntensity, PT_TONE_MAPPING_SDR };
    std::vector<uint8_t> out(static_cast<size_t>(m_Width) * m_Height * 4);
    Check(pt_postprocess(m_Context, &params, std::max(1u, m_TotalSamples), PT_OUTPUT_RGBA8_SRGB, out.data(), out.size()),
          "pt_postprocess");
    return out;
}

void HeadlessRenderer::SavePng(const std::string &path)
{
    const std::vector<uint8_t> out = PostProcessSrgb8();
    if (stbi_write_png(path.c_str(), m_Width, m_Height, 4, out.data(), 0) == 0)
        throw error(std::format("Could not write {}", path));
}

/* OutputSaver.cpp:237-242: quality argument 0 (stb's default), four components */
void HeadlessRenderer::SaveJpg(const std::string &path)
{
    const std::vector<uint8_t> out = PostProcessSrgb8();
    if (stbi_write_jpg(path.c_str(), m_Width, m_Height, 4, out.data(), 0) == 0)
        throw error(std::format("Could not write {}", path));
}

void HeadlessRenderer::SaveTga(const std::string &path)
{
    const std::vector<uint8_t> out = PostProcessSrgb8();
    if (stbi_write_tga(path.c_str(), m_Width, m_Height, 4, out.data()) == 0)
        throw error(std::format("Could not write {}", path));
}

void HeadlessRenderer::SaveHdr(const std::string &path)
{
    const pt_postprocess_params params = { m_PostProcess.Exposure, m_PostProcess.BloomThreshold,
                                           m_PostProcess.BloomIntensity, PT_TONE_MAPPING_HDR };
    std::vector<float> out(static_cast<size_t>(m_Width) * m_Height * 4);
    Check(pt_postprocess(m_Context, &params, std::max(1u, m_TotalSamples), PT_OUTPUT_RGBAF32, out.data(),
                         out.size() * sizeof(float)),
          "pt_postprocess");
    if (stbi_write_hdr(path.c_str(), m_Width, m_Height, 4, out.data()) == 0)
        throw error(std::format("Could not write {}", path));
}

pt_stats HeadlessRenderer::GetStats()
{
    pt_stats stats = {};
    Check(pt_get_stats(m_Context, &stats), "pt_get_stats");
    return stats;
}

}
