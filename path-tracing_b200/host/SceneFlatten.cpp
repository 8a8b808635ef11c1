#include "HeadlessRenderer.h"

#include <cstring>
#include <format>

#include "Core/Core.h"

#include "TextureImporter.h"

namespace PathTracing
{

static_assert(sizeof(pt_vertex) == sizeof(Shaders::Vertex));
static_assert(sizeof(pt_material_mr) == sizeof(Shaders::MetallicRoughnessMaterial));
static_assert(sizeof(pt_material_sg) == sizeof(Shaders::SpecularGlossinessMaterial));
static_assert(sizeof(pt_material_phong) == sizeof(Shaders::PhongMaterial));
static_assert(sizeof(pt_point_light) == sizeof(Shaders::PointLight));
static_assert(sizeof(pt_directional_light) == sizeof(Shaders::DirectionalLight));
static_assert(PT_SCENE_TEXTURE_OFFSET == Shaders::SceneTextureOffset);
static_assert(PT_MAX_LIGHT_COUNT == Shaders::MaxLightCount);

namespace
{

/* TextureUploader::GetImageFormat's colour-space rule (TextureUploader.cpp:571-595) */
bool IsSrgbTexture(TextureType type)
{
    return type == TextureType::Color || type == TextureType::Specular || type == TextureType::Emisive ||
           type == TextureType::Skybox;
}

pt_texture_desc LoadTexture(const TextureInfo &info, std::vector<std::byte> &pixels)
{
    TextureData data = TextureImporter::LoadTextureData(info);
    pixels.assign(data.begin(), data.end());
    TextureImporter::ReleaseTextureData(info, data);

    pt_texture_desc desc = {};
    desc.width = info.Width;
    desc.height = info.Height;
    switch (info.Format)
    {
    case TextureFormat::RGBAU8: desc.format = PT_TEXTURE_RGBA8; break;
    case TextureFormat::RGBAF32: desc.format = PT_TEXTURE_RGBAF32; break;
    /* .dds files: the gli loader returns every stored level, level 0 first (TextureImporter.cpp:311-343) */
    case TextureFormat::BC1: desc.format = PT_TEXTURE_BC1; break;
    case TextureFormat::BC3: desc.format = PT_TEXTURE_BC3; break;
    case TextureFormat::BC5: desc.format = PT_TEXTURE_BC5; break;
    default: throw error(std::format("Texture {}: unsupported texture format", info.Name));
    }
    desc.srgb = (info.Format == TextureFormat::RGBAU8 || info.Format == TextureFormat::BC1) && IsSrgbTexture(info.Type);
    desc.levels = info.Levels;
    desc.pixels = pixels.data();
    return desc;
}

}

std::unique_ptr<FlattenedScene> FlattenScene(const Scene &scene, bool loadTextures)
{
    auto flat = std::make_unique<FlattenedScene>();
    pt_scene_desc &desc = flat->Desc;

    const auto vertices = scene.GetVertices();
    desc.vertices = reinterpret_cast<const pt_vertex *>(vertices.data());
    desc.vertex_count = vertices.size();

    const auto indices = scene.GetIndices();
    desc.indices = indices.data();
    desc.index_count = indices.size();

    /* glm::mat3x4 = 3 columns of 4 = the three rows of the 3x4 matrix (Scene.h:306,312) */
    const auto transforms = scene.GetTransforms();
    flat->Transforms.resize(transforms.size() * 12);
    std::memcpy(flat->Transforms.data(), transforms.data(), flat->Transforms.size() * sizeof(float));
    desc.transforms = flat->Transforms.data();
    desc.transform_count = static_cast<uint32_t>(transforms.size());

    /* The reference keeps non-animated geometries first and appends one entry per animated mesh
     * instance, addressed through geometryIndexMap (Renderer.cpp:333-372).  The C ABI takes the scene's
     * geometry list as it is plus the IsAnimated flags: an animated geometry's offsets address the
     * animated vertex / index buffers, and the core skins it (skinning.comp) before baking. */
    const auto geometries = scene.GetGeometries();
    std::vector<uint32_t> geometryIndexMap(geometries.size(), 0);
    bool anyAnimated = false;
    for (size_t i = 0; i < geometries.size(); i++)
    {
        const Geometry &geometry = geometries[i];
        geometryIndexMap[i] = static_cast<uint32_t>(flat->Geometries.size());
        flat->Geometries.push_back(pt_geometry { geometry.VertexOffset, geometry.VertexLength, geometry.IndexOffset,
                                                 geometry.IndexLength, geometry.IsOpaque ? 1u : 0u });
        flat->GeometryIsAnimated.push_back(geometry.IsAnimated ? 1u : 0u);
        anyAnimated |= geometry.IsAnimated;
    }
    desc.geometries = flat->Geometries.data();
    desc.geometry_count = static_cast<uint32_t>(flat->Geometries.size());
    if (anyAnimated)
    {
        static_assert(sizeof(pt_animated_vertex) == sizeof(Shaders::AnimatedVertex));
        const auto animatedVertices = scene.GetAnimatedVertices();
        const auto animatedIndices = scene.GetAnimatedIndices();
        const auto bones = scene.GetBoneTransforms();
        desc.geometry_is_animated = flat->GeometryIsAnimated.data();
        desc.animated_vertices = reinterpret_cast<const pt_animated_vertex *>(animatedVertices.data());
        desc.animated_vertex_count = animatedVertices.size();
        desc.animated_indices = animatedIndices.data();
        desc.animated_index_count = animatedIndices.size();
        desc.bone_transforms = reinterpret_cast<const float *>(bones.data());
        desc.bone_count = static_cast<uint32_t>(bones.size());
    }

    /* one record per (model, mesh) in model order (Renderer.cpp:381-399) */
    const auto models = scene.GetModels();
    for (const Model &model : models)
    {
        flat->Models.push_back(pt_model { model.MeshOffset, static_cast<uint32_t>(model.Meshes.size()) });
        for (const Mesh &mesh : model.Meshes)
            flat->MeshRecords.push_back(
                pt_mesh_record { geometryIndexMap[mesh.GeometryIndex], mesh.MaterialIndex, mesh.TransformBufferOffset }
            );
    }
    desc.models = flat->Models.data();
    desc.model_count = static_cast<uint32_t>(flat->Models.size());
    desc.mesh_records = flat->MeshRecords.data();
    desc.mesh_record_count = static_cast<uint32_t>(flat->MeshRecords.size());

    /* TLAS instances (AccelerationStructure.cpp:268-275): first 12 floats of the row-vector mat4 */
    for (const ModelInstance &instance : scene.GetModelInstances())
    {
        pt_instance out = {};
        std::memcpy(out.transform, &instance.Transform, sizeof(out.transform));
        out.model_index = instance.ModelIndex;
        flat->Instances.push_back(out);
    }
    desc.instances = flat->Instances.data();
    desc.instance_count = static_cast<uint32_t>(flat->Instances.size());

    const auto mr = scene.GetMetallicRoughnessMaterials();
    desc.mr_materials = reinterpret_cast<const pt_material_mr *>(mr.data());
    desc.mr_material_count = static_cast<uint32_t>(mr.size());
    const auto sg = scene.GetSpecularGlossinessMaterials();
    desc.sg_materials = reinterpret_cast<const pt_material_sg *>(sg.data());
    desc.sg_material_count = static_cast<uint32_t>(sg.size());
    const auto phong = scene.GetPhongMaterials();
    desc.phong_materials = reinterpret_cast<const pt_material_phong *>(phong.data());
    desc.phong_material_count = static_cast<uint32_t>(phong.size());

    if (loadTextures)
    {
        const auto textures = scene.GetTextures();
        flat->TexturePixels.resize(textures.size());
        for (size_t i = 0; i < textures.size(); i++)
            flat->Textures.push_back(LoadTexture(textures[i], flat->TexturePixels[i]));
        desc.textures = flat->Textures.data();
        desc.texture_count = static_cast<uint32_t>(flat->Textures.size());

        if (const Skybox2D *skybox = std::get_if<Skybox2D>(&scene.GetSkybox()))
        {
            flat->Skybox2D = LoadTexture(skybox->Content, flat->SkyboxPixels);
            desc.skybox_2d = &flat->Skybox2D;
        }
        else if (const SkyboxCube *cube = std::get_if<SkyboxCube>(&scene.GetSkybox()))
        {
            /* layer order of TextureUploader::UploadSkyboxBlocking (TextureUploader.cpp:232-236) */
            const TextureInfo *faces[6] = { &cube->Front, &cube->Back, &cube->Up, &cube->Down, &cube->Left, &cube->Right };
            for (int i = 0; i < 6; i++)
                flat->SkyboxCube[i] = LoadTexture(*faces[i], flat->SkyboxCubePixels[i]);
            desc.skybox_cube = flat->SkyboxCube;
        }
    }

    /* light UBO content (Renderer.cpp:1719-1726) */
    const auto pointLights = scene.GetPointLights();
    desc.point_lights = reinterpret_cast<const pt_point_light *>(pointLights.data());
    desc.point_light_count = static_cast<uint32_t>(pointLights.size());
    std::memcpy(&desc.directional_light, &scene.GetDirectionalLight(), sizeof(pt_directional_light));

    return flat;
}

}
