/*
 * HeadlessRenderer — host-side drop-in for the reference's `Renderer` facade
 * (Path-Tracing/Renderer/Renderer.h:42-85) on top of the C ABI in include/pt_core.h.
 *
 * It is compiled against the UNMODIFIED reference headers (Scene.h, Core/Camera.h,
 * Shaders/ShaderTypes.incl) and keeps the reference's verbs and settings structs, so
 * Application::Run's three calls (Path-Tracing/Application.cpp:337-351)
 *     Renderer::UpdateSceneData(scene, updated); Renderer::OnUpdate(dt); Renderer::Render();
 * map 1:1.  Instead of a swapchain it owns a float4 accumulation buffer on the GPU
 * and hands the sum image back to the host.
 *
 * Errors from the C ABI are rethrown as PathTracing::error, the reference's own
 * exception type (Path-Tracing/Core/Core.cpp:82-90).
 */
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "Scene.h"

#include "pt_core.h"

namespace PathTracing
{

/* Owns the POD arrays a pt_scene_desc points into. */
struct FlattenedScene
{
    std::vector<float> Transforms; /* 12 per transform */
    std::vector<pt_geometry> Geometries;
    std::vector<uint32_t> GeometryIsAnimated;
    std::vector<pt_mesh_record> MeshRecords;
    std::vector<pt_model> Models;
    std::vector<pt_instance> Instances;
    std::vector<pt_texture_desc> Textures;
    std::vector<std::vector<std::byte>> TexturePixels;
    pt_texture_desc Skybox2D = {};
    std::vector<std::byte> SkyboxPixels;
    pt_texture_desc SkyboxCube[6] = {};
    std::vector<std::byte> SkyboxCubePixels[6];
    pt_scene_desc Desc = {};
};

/* What Renderer::UpdateSceneData pulls out of a Scene (Renderer.cpp:257-399), as PODs.
 * Scene::Update must have run once so that instance transforms are final (Scene.cpp:65-70).
 * Texture pixels are decoded with the reference's own TextureImporter. */
std::unique_ptr<FlattenedScene> FlattenScene(const Scene &scene, bool loadTextures = true);

class HeadlessRenderer
{
public:
    /* Renderer::PathTracingSettings, Renderer.h:55-60 */
    struct PathTracingSettings
    {
        uint32_t BounceCount = 4;
        float LensRadius = 0.0f;
        float FocalDistance = 10.0f;
    };

    /* Renderer::PostProcessSettings, Renderer.h:68-73 */
    struct PostProcessSettings
    {
        float Exposure = 1.0f;
        float BloomThreshold = 1.0f;
        float BloomIntensity = 0.1f;
    };

    /* Renderer::SetDebugRaytracingPipeline's PipelineConfig<4> (Renderer.h:30, 58): the four specialisation
     * constants of the debug pipeline in constant-id order (Debug/DebugShaderTypes.incl:13-16) */
    struct DebugRaytracingPipelineConfig
    {
        uint32_t RenderMode = 0;
        uint32_t RaygenFlags = 0;
        uint32_t MissFlags = 0;
        uint32_t HitGroupFlags = 0;
    };

    explicit HeadlessRenderer(int cudaDevice = 0); /* Renderer::Init     */
    ~HeadlessRenderer();                           /* Renderer::Shutdown */

    HeadlessRenderer(const HeadlessRenderer &) = delete;
    HeadlessRenderer &operator=(const HeadlessRenderer &) = delete;

    /* Renderer::UpdateSceneData (Renderer.cpp:238-439) */
    void UpdateSceneData(const std::shared_ptr<Scene> &scene, bool updated);
    /* Renderer::OnResize (Renderer.cpp:801-808 resets accumulation) */
    void OnResize(uint32_t width, uint32_t height);
    void SetSettings(const PathTracingSettings &settings);
    void SetSettings(const PostProcessSettings &settings);

    /* Renderer::RenderSettings::SamplesPerFrame / the adaptive s_SamplesPerFrame (Renderer.cpp:1631-1657): samples one
     * frame = one vkCmdTraceRaysKHR renders per pixel, on one rng stream (raygen.rgen:36-118).  Default 1, the value of
     * the reference's Profile / Debug builds (Core/Config.h:34-36). */
    void SetSamplesPerFrame(uint32_t samplesPerFrame);
    /* Renderer::Render (Renderer.cpp:1659-1808), `frames` times: every frame adds SamplesPerFrame samples */
    void Render(uint32_t frames = 1);
    /* Renderer::SetDebugRaytracingPipeline + one frame of that pipeline: width*height RGBA floats */
    void SetDebugRaytracingPipeline(const DebugRaytracingPipelineConfig &config);
    [[nodiscard]] std::vector<float> RenderDebug();

    [[nodiscard]] uint32_t GetTotalSamples() const { return m_TotalSamples; }
    /* raw sum image, RGBA float, width*height*4 */
    [[nodiscard]] std::vector<float> ReadAccumulation();
    /* exposure + bloom + composition + SDR tone mapping + sRGB8 (pt_postprocess), written with stb like OutputSaver */
    void SavePng(const std::string &path);
    /* OutputFormat::Jpg / Tga (OutputSaver.cpp:237-242): the same 8-bit sRGB pixels through stb's writers */
    void SaveJpg(const std::string &path);
    void SaveTga(const std::string &path);
    /* OutputFormat::Hdr: same chain without the tone curve, as Radiance .hdr */
    void SaveHdr(const std::string &path);

    [[nodiscard]] pt_stats GetStats();

private:
    void Check(pt_status status, const char *what);
    pt_render_params MakeRenderParams();
    void ResetAccumulation();
    std::vector<uint8_t> PostProcessSrgb8();

    pt_context *m_Context = nullptr;
    std::shared_ptr<Scene> m_Scene;
    uint32_t m_Width = 0, m_Height = 0;
    uint32_t m_TotalSamples = 0;
    uint32_t m_SamplesPerFrame = 1;
    PathTracingSettings m_PathTracing;
    PostProcessSettings m_PostProcess;
    DebugRaytracingPipelineConfig m_Debug;
};

}
