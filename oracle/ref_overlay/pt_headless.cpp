/*
 * pt_headless — the reference's host code driving the B200 core: SceneManager loads a scene
 * with the reference's own loaders, HeadlessRenderer (the Renderer stand-in) uploads it through
 * the C ABI, renders `spp` samples and writes a PNG, i.e. the reference's "Render" button flow
 * (Path-Tracing/UserInterface.cpp:1074-1092) without a window.
 *
 *   pt_headless <group> <scene> <width> <height> <spp> <bounces> <out>
 *   pt_headless --file <model.gltf|.glb|.obj> <width> <height> <spp> <bounces> <out>   (assimp overlay)
 * <out>: .png / .jpg / .tga / .hdr like OutputSaver's formats (Renderer/OutputSaver.h:16-19), or .f32 = the raw float4
 * accumulation (sum) image, width*height*16 bytes — what tests/test_gpu_shim.py compares bit for bit.
 * PT_SAMPLES_PER_FRAME=n renders frames of n samples (RenderSettings::SamplesPerFrame); <spp> must be a multiple.
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "Core/Core.h"

#include "HeadlessRenderer.h"
#include "SceneImporter.h"
#include "SceneManager.h"

using namespace PathTracing;

int main(int argc, char **argv)
{
    if (argc != 8)
    {
        std::fprintf(stderr, "usage: pt_headless <group> <scene> <width> <height> <spp> <bounces> <out.png>\n");
        return 2;
    }
    try
    {
        const bool fromFile = std::string(argv[1]) == "--file";
        std::shared_ptr<Scene> scene;
        if (fromFile)
        {
            SceneImporter::Init();
            SceneBuilder builder;
            SceneImporter::AddFile(builder, argv[2]);
            scene = builder.CreateSceneShared(std::filesystem::path(argv[2]).stem().string());
        if (scene->GetSceneCamerasCount() > 0)
            scene->SetActiveCamera(0); /* the file's own first camera instead of the input camera */
        }
        else
        {
            SceneManager::Init();
            if (std::string(argv[1]) != "Test Scenes" || std::string(argv[2]) != "Default")
                SceneManager::SetActiveScene(argv[1], argv[2]);
            scene = SceneManager::GetActiveScene();
        }
        InputCamera::DisableInput();

        HeadlessRenderer renderer(0);
        renderer.OnResize(std::atoi(argv[3]), std::atoi(argv[4]));
        renderer.SetSettings(HeadlessRenderer::PathTracingSettings { .BounceCount = (uint32_t)std::atoi(argv[6]) });

        const uint32_t spp = std::atoi(argv[5]);
        const bool updated = scene->Update(0.0f);
        renderer.UpdateSceneData(scene, updated);
        uint32_t perFrame = 1;
        if (const char *e = std::getenv("PT_SAMPLES_PER_FRAME"))
            perFrame = std::max(1, std::atoi(e));
        renderer.SetSamplesPerFrame(perFrame);
        renderer.Render(spp / perFrame);
        const std::string out = argv[7];
        const std::string ext = out.size() >= 4 ? out.substr(out.size() - 4) : "";
        if (ext == ".jpg")
            renderer.SaveJpg(out);
        else if (ext == ".tga")
            renderer.SaveTga(out);
        else if (ext == ".hdr")
            renderer.SaveHdr(out);
        else if (ext == ".f32")
        {
            const std::vector<float> acc = renderer.ReadAccumulation();
            FILE *f = std::fopen(out.c_str(), "wb");
            if (!f || std::fwrite(acc.data(), sizeof(float), acc.size(), f) != acc.size())
                throw error("Could not write " + out);
            std::fclose(f);
        }
        else
            renderer.SavePng(out);

        const pt_stats stats = renderer.GetStats();
        std::printf(
            "%u spp, %.3f ms, %.1f Mrays/s (%llu closest + %llu shadow), %llu tris, BVH build %.3f ms\n",
            renderer.GetTotalSamples(), stats.last_render_ms,
            (stats.rays_closest + stats.rays_shadow) / (stats.last_render_ms * 1e3),
            (unsigned long long)stats.rays_closest, (unsigned long long)stats.rays_shadow,
            (unsigned long long)stats.triangle_count, stats.bvh_build_ms
        );
        if (!fromFile)
            SceneManager::Shutdown();
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
