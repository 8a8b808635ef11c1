/*
 * scene_dump — runs the reference's OWN host code (SceneManager -> ExampleScenes ->
 * SceneBuilder -> Scene::Update -> Camera) for a built-in scene, flattens it with the product's
 * FlattenScene() and writes the PODs as a chunked binary file that tests/golden/make_default_scene.py
 * turns into the committed fixture.  Built only where /root/reference exists.
 *
 *   scene_dump <group> <scene> <width> <height> <out.ptscene>
 *   scene_dump --file <model.gltf|.glb|.obj> <width> <height> <out.ptscene>     (needs the assimp overlay:
 *       the reference's own SceneImporter::AddFile, like SceneManager's file loaders, SceneManager.cpp:43-52)
 */
#include <cstdio>
#include <cstring>
#include <fstream>

#include "Core/Core.h"

#include "HeadlessRenderer.h"
#include "SceneImporter.h"
#include "SceneManager.h"

using namespace PathTracing;

static void WriteChunk(std::ofstream &out, const char *name, const void *data, uint64_t bytes)
{
    char tag[24] = {};
    std::strncpy(tag, name, sizeof(tag) - 1);
    out.write(tag, sizeof(tag));
    out.write(reinterpret_cast<const char *>(&bytes), sizeof(bytes));
    out.write(reinterpret_cast<const char *>(data), static_cast<std::streamsize>(bytes));
}

int main(int argc, char **argv)
{
    if (argc != 6)
    {
        std::fprintf(stderr, "usage: scene_dump <group> <scene> <width> <height> <out.ptscene>\n");
        return 2;
    }
    const uint32_t width = std::atoi(argv[3]), height = std::atoi(argv[4]);

    std::shared_ptr<Scene> scene;
    if (std::string(argv[1]) == "--file")
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
        SceneManager::Init(); /* loads "Test Scenes"/"Default" like Application::Init */
        if (std::string(argv[1]) != "Test Scenes" || std::string(argv[2]) != "Default")
            SceneManager::SetActiveScene(argv[1], argv[2]);
        scene = SceneManager::GetActiveScene();
    }
    InputCamera::DisableInput();
    scene->Update(0.0f); /* instance transforms become final (Scene.cpp:65-70) */

    const auto flat = FlattenScene(*scene);
    const pt_scene_desc &d = flat->Desc;

    Camera &camera = scene->GetActiveCamera();
    camera.OnResize(width, height);
    const glm::mat4 invView = camera.GetInvViewMatrix();
    const glm::mat4 invProj = camera.GetInvProjectionMatrix();

    std::ofstream out(argv[5], std::ios::binary);
    out.write("PTSCENE1", 8);
    WriteChunk(out, "vertices", d.vertices, d.vertex_count * sizeof(pt_vertex));
    WriteChunk(out, "indices", d.indices, d.index_count * sizeof(uint32_t));
    WriteChunk(out, "transforms", d.transforms, d.transform_count * 48ull);
    WriteChunk(out, "geometries", d.geometries, d.geometry_count * sizeof(pt_geometry));
    WriteChunk(out, "mesh_records", d.mesh_records, d.mesh_record_count * sizeof(pt_mesh_record));
    WriteChunk(out, "models", d.models, d.model_count * sizeof(pt_model));
    WriteChunk(out, "instances", d.instances, d.instance_count * sizeof(pt_instance));
    WriteChunk(out, "mr_materials", d.mr_materials, d.mr_material_count * sizeof(pt_material_mr));
    WriteChunk(out, "sg_materials", d.sg_materials, d.sg_material_count * sizeof(pt_material_sg));
    WriteChunk(out, "phong_materials", d.phong_materials, d.phong_material_count * sizeof(pt_material_phong));
    WriteChunk(out, "point_lights", d.point_lights, d.point_light_count * sizeof(pt_point_light));
    WriteChunk(out, "directional_light", &d.directional_light, sizeof(pt_directional_light));
    if (d.geometry_is_animated)
    {
        WriteChunk(out, "geometry_is_animated", d.geometry_is_animated, d.geometry_count * sizeof(uint32_t));
        WriteChunk(out, "animated_vertices", d.animated_vertices, d.animated_vertex_count * sizeof(pt_animated_vertex));
        WriteChunk(out, "animated_indices", d.animated_indices, d.animated_index_count * sizeof(uint32_t));
        WriteChunk(out, "bone_transforms", d.bone_transforms, d.bone_count * 48ull);
    }
    for (uint32_t i = 0; i < d.texture_count; i++)
    {
        const pt_texture_desc &t = d.textures[i];
        const uint32_t info[4] = { t.width, t.height, t.format, t.srgb };
        WriteChunk(out, "texture_info", info, sizeof(info));
        if (t.format >= PT_TEXTURE_BC1)
            throw error("scene_dump: block-compressed textures are not written to fixtures");
        const uint64_t bpp = t.format == PT_TEXTURE_RGBAF32 ? 16 : 4;
        WriteChunk(out, "texture_pixels", t.pixels, bpp * t.width * t.height);
    }
    const uint32_t extent[2] = { width, height };
    WriteChunk(out, "camera_extent", extent, sizeof(extent));
    WriteChunk(out, "view_inverse", &invView, sizeof(invView));
    WriteChunk(out, "proj_inverse", &invProj, sizeof(invProj));
    const uint32_t flags[2] = { std::holds_alternative<Skybox2D>(scene->GetSkybox()) ? 1u : 0u,
                                scene->HasDxNormalTextures() ? 1u : 0u };
    WriteChunk(out, "flags", flags, sizeof(flags));
    out.close();

    std::printf(
        "%s/%s: %llu vertices, %llu indices, %u geometries, %u transforms, %u mesh records, %u models, "
        "%u instances, %u MR materials, %u textures, %u point lights\n",
        argv[1], argv[2], (unsigned long long)d.vertex_count, (unsigned long long)d.index_count, d.geometry_count,
        d.transform_count, d.mesh_record_count, d.model_count, d.instance_count, d.mr_material_count,
        d.texture_count, d.point_light_count
    );
    if (std::string(argv[1]) != "--file")
        SceneManager::Shutdown();
    return 0;
}
