/*
 * glsl_rt_api.h — C ABI of oracle/_ref/libglsl_ref.so (third runtime header, included after the
 * generated stages; see glsl_rt.h).  TEST INFRASTRUCTURE: used only by tests/ to pin the CPU oracle
 * (oracle/pt_oracle.cpp) to the reference's own shader text.
 *
 *   glr_test_shading   per-function probes, same modes and record layouts as pto_test_shading /
 *                      pt_test_shading (include/pt_core.h), each calling the function the GLSL defines;
 *   glr_closest_hit    closestHit.rchit main() on given hits + payloads;
 *   glr_render         raygen.rgen main() per pixel and frame, traceRayEXT dispatching to
 *                      anyhit.rahit / occlusionAnyhit.rahit / closestHit.rchit / miss.rmiss /
 *                      occlusion.rmiss main()s.
 */
#pragma once

#include <atomic>
#include <stdexcept>
#include <thread>
#include <vector>

namespace glslref
{

struct Scene
{
    std::vector<pt_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<float> transformData;
    std::vector<pt_geometry> geometryDescs;
    std::vector<Geometry> geometryTable;
    std::vector<pt_mesh_record> meshRecords;
    std::vector<pt_model> models;
    std::vector<pt_instance> instances;
    std::vector<pt_material_mr> mr;
    std::vector<pt_material_sg> sg;
    std::vector<pt_material_phong> phong;
    std::vector<pt_point_light> pointLights;
    pt_directional_light directional;
    glr_callbacks cb;
};

/* ---- sampler / image / traversal entry points of the stages ---- */

inline vec4 sampleThroughCallback(uint32_t slot, vec2 uv, vec2 dx, vec2 dy, int useGrad)
{
    const float in6[6] = { uv.x, uv.y, dx.x, dx.y, dy.x, dy.y };
    float out4[4] = { 0, 0, 0, 0 };
    tls_scene->cb.texture(tls_scene->cb.user, slot, in6, out4, 1, useGrad);
    return vec4(out4[0], out4[1], out4[2], out4[3]);
}
/* texture() outside a fragment stage has no implicit derivatives: level 0 */
vec4 texture(sampler2D s, vec2 uv)
{
    if (s.slot == 0xffffffffu) /* skybox2D */
    {
        const float in3[3] = { uv.x, uv.y, 0.0f };
        float out4[4] = { 0, 0, 0, 0 };
        tls_scene->cb.sky(tls_scene->cb.user, 0, in3, out4);
        return vec4(out4[0], out4[1], out4[2], out4[3]);
    }
    return sampleThroughCallback(s.slot, uv, vec2(0.0f), vec2(0.0f), 0);
}
vec4 textureGrad(sampler2D s, vec2 uv, vec2 dPdx, vec2 dPdy) { return sampleThroughCallback(s.slot, uv, dPdx, dPdy, 1); }
vec4 texture(samplerCube, vec3 dir)
{
    const float in3[3] = { dir.x, dir.y, dir.z };
    float out4[4] = { 0, 0, 0, 0 };
    tls_scene->cb.sky(tls_scene->cb.user, 1, in3, out4);
    return vec4(out4[0], out4[1], out4[2], out4[3]);
}
vec4 imageLoad(image2D, ivec2 p)
{
    const float *px = tls_image + 4 * ((size_t)p.y * gl_LaunchSizeEXT.x + (size_t)p.x);
    return vec4(px[0], px[1], px[2], px[3]);
}
void imageStore(image2D, ivec2 p, vec4 v)
{
    float *px = tls_image + 4 * ((size_t)p.y * gl_LaunchSizeEXT.x + (size_t)p.x);
    px[0] = v.x, px[1] = v.y, px[2] = v.z, px[3] = v.w;
}

/* shader record + built-ins of one (candidate) hit */
inline void bindHit(const Scene &s, uint32_t instance, uint32_t geometry, uint32_t primitive, float t, float b1, float b2)
{
    const pt_instance &inst = s.instances[instance];
    /* instanceShaderBindingTableRecordOffset = 2 * MeshOffset, geometry index within the BLAS,
     * stride 2 (AccelerationStructure.cpp:268-275, raygen.rgen:31,68) */
    const pt_mesh_record &rec = s.meshRecords[s.models[inst.model_index].mesh_offset + geometry];
    sbt.GeometryIndex = rec.geometry_index;
    sbt.MaterialId = rec.material_id;
    sbt.TransformIndex = rec.transform_index;
    gl_PrimitiveID = (int)primitive;
    gl_GeometryIndexEXT = (int)geometry;
    gl_InstanceID = (int)instance;
    gl_RayTmaxEXT = t;
    attribs = vec3(b1, b2, 0.0f);
    const float *m = inst.transform;
    gl_ObjectToWorld3x4EXT = mat3x4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11]);
}

struct AnyHitCtx
{
    const Scene *scene;
    bool occlusion;
};

inline int32_t anyHitTrampoline(void *ctx, uint32_t instance, uint32_t geometry, uint32_t primitive, float t, float b1,
                                float b2)
{
    const AnyHitCtx *c = (const AnyHitCtx *)ctx;
    bindHit(*c->scene, instance, geometry, primitive, t, b1, b2);
    tls_ignore = false;
    if (c->occlusion)
        occ_rahit::main_();
    else if (tls_debug_pipeline)
        dbg_rahit::main_();
    else
        rahit::main_();
    return tls_ignore ? 0 : 1;
}

void traceRayEXT(accelerationStructureEXT, uint rayFlags, uint, uint sbtRecordOffset, uint, uint missIndex, vec3 origin,
                 float tmin, vec3 direction, float tmax, int)
{
    const Scene &s = *tls_scene;
    if (++tls_trace_calls > (1ull << 24))
        throw std::runtime_error("raygen.rgen: NaN/Inf restart loop does not terminate");
    /* built-ins and the shader record belong to ONE invocation: a stage that traces (debugClosestHit.rchit's shadow rays)
     * must find its own again afterwards */
    struct Saved
    {
        SBTBuffer sbt_;
        int prim, geom, inst;
        float tmaxv;
        vec3 attr, org, dir;
        mat3x4 o2w;
        Saved() : sbt_(sbt), prim(gl_PrimitiveID), geom(gl_GeometryIndexEXT), inst(gl_InstanceID), tmaxv(gl_RayTmaxEXT), attr(attribs),
                  org(gl_WorldRayOriginEXT), dir(gl_WorldRayDirectionEXT), o2w(gl_ObjectToWorld3x4EXT) {}
        ~Saved()
        {
            sbt = sbt_, gl_PrimitiveID = prim, gl_GeometryIndexEXT = geom, gl_InstanceID = inst, gl_RayTmaxEXT = tmaxv, attribs = attr;
            gl_WorldRayOriginEXT = org, gl_WorldRayDirectionEXT = dir, gl_ObjectToWorld3x4EXT = o2w;
        }
    } saved;
    gl_WorldRayOriginEXT = origin;
    gl_WorldRayDirectionEXT = direction;
    const bool occlusion = sbtRecordOffset == OcclusionRayHitGroupIndex;
    AnyHitCtx ctx { &s, occlusion };
    pt_hit hit;
    const float o[3] = { origin.x, origin.y, origin.z }, d[3] = { direction.x, direction.y, direction.z };
    const uint terminate = (rayFlags & gl_RayFlagsTerminateOnFirstHitEXT) ? 1u : 0u;
    const uint flags = ((rayFlags & gl_RayFlagsOpaqueEXT) ? 1u : 0u) | ((rayFlags & gl_RayFlagsCullBackFacingTrianglesEXT) ? 2u : 0u);
    const int32_t found = flags ? s.cb.trace_flags(s.cb.user, o, d, tmin, tmax, terminate, flags, anyHitTrampoline, &ctx, &hit)
                                : s.cb.trace(s.cb.user, o, d, tmin, tmax, terminate, anyHitTrampoline, &ctx, &hit);
    if (found)
    {
        /* the occlusion hit group has no closest-hit shader (Renderer.cpp pipeline: any-hit only) */
        if (!occlusion)
        {
            bindHit(s, hit.instance, hit.geometry, hit.primitive, hit.t, hit.u, hit.v);
            if (tls_debug_pipeline)
                dbg_rchit::main_();
            else
                rchit::main_();
        }
    }
    else if (missIndex == OcclusionRayMissGroupIndex)
        occ_rmiss::main_(); /* both pipelines (Renderer.cpp:558, 589) */
    else if (tls_debug_pipeline)
        dbg_rmiss::main_();
    else
        rmiss::main_();
}

inline void bindScene(const Scene &s, const pt_render_params &p)
{
    tls_scene = &s;
    transforms = (const mat3x4 *)s.transformData.data();
    geometries = s.geometryTable.data();
    metallicRoughnessMaterials = (const MetallicRoughnessMaterial *)s.mr.data();
    specularGlossinessMaterials = (const SpecularGlossinessMaterial *)s.sg.data();
    phongMaterials = (const PhongMaterial *)s.phong.data();
    u_LightCount = (uint)s.pointLights.size();
    std::memcpy(&u_DirectionalLight, &s.directional, sizeof(DirectionalLight));
    u_Lights = (const PointLight *)s.pointLights.data();
    skybox2D = sampler2D { 0xffffffffu };
    s_HitFlags = p.hit_flags;
    s_MissFlags = p.miss_flags;
    std::memcpy(&mainUniform.MainCamera.ViewInverse, p.view_inverse, 64);
    std::memcpy(&mainUniform.MainCamera.ProjInverse, p.proj_inverse, 64);
    mainUniform.BounceCount = p.bounce_count;
    mainUniform.LensRadius = p.lens_radius;
    mainUniform.FocalDistance = p.focal_distance;
}

inline MaterialSample materialFromFloats(const float *f)
{
    MaterialSample m;
    m.EmissiveColor = vec3(f[0], f[1], f[2]);
    m.Color = vec3(f[3], f[4], f[5]);
    m.Normal = vec3(f[6], f[7], f[8]);
    m.Roughness = f[9];
    m.Metalness = f[10];
    m.Transmission = f[11];
    m.Eta = f[12];
    m.AttenuationColor = vec3(f[13], f[14], f[15]);
    m.AttenuationDistance = f[16];
    return m;
}

inline Vertex vertexPUN(const float *a) /* position, uv, normal */
{
    Vertex v;
    v.Position = vec3(a[0], a[1], a[2]);
    v.TexCoords = vec2(a[3], a[4]);
    v.Normal = vec3(a[5], a[6], a[7]);
    v.Tangent = vec3(0.0f);
    v.Bitangent = vec3(0.0f);
    return v;
}

} // namespace glslref

extern "C" {

struct glr_scene
{
    glslref::Scene s;
};

PT_API glr_scene *glr_scene_create(const pt_scene_desc *d, const glr_callbacks *cb)
{
    if (!d || !cb)
        return nullptr;
    if (d->geometry_is_animated)
        for (uint32_t g = 0; g < d->geometry_count; g++)
            if (d->geometry_is_animated[g])
                return nullptr; /* skinning.comp is not part of the ray-tracing stages */
    glr_scene *h = new glr_scene();
    glslref::Scene &s = h->s;
    s.vertices.assign(d->vertices, d->vertices + d->vertex_count);
    s.indices.assign(d->indices, d->indices + d->index_count);
    s.transformData.assign(d->transforms, d->transforms + 12 * (size_t)d->transform_count);
    s.geometryDescs.assign(d->geometries, d->geometries + d->geometry_count);
    for (const pt_geometry &g : s.geometryDescs)
    {
        /* Renderer.cpp:338-350: buffer base + offset; indices stay relative to the first vertex */
        glslref::Geometry e;
        e.Vertices.v = (glm::vec2 *)(s.vertices.data() + g.vertex_offset);
        e.Indices.v = s.indices.data() + g.index_offset;
        s.geometryTable.push_back(e);
    }
    s.meshRecords.assign(d->mesh_records, d->mesh_records + d->mesh_record_count);
    s.models.assign(d->models, d->models + d->model_count);
    s.instances.assign(d->instances, d->instances + d->instance_count);
    if (d->mr_material_count)
        s.mr.assign(d->mr_materials, d->mr_materials + d->mr_material_count);
    if (d->sg_material_count)
        s.sg.assign(d->sg_materials, d->sg_materials + d->sg_material_count);
    if (d->phong_material_count)
        s.phong.assign(d->phong_materials, d->phong_materials + d->phong_material_count);
    if (d->point_light_count)
        s.pointLights.assign(d->point_lights, d->point_lights + d->point_light_count);
    s.directional = d->directional_light;
    s.cb = *cb;
    return h;
}

PT_API void glr_scene_destroy(glr_scene *s) { delete s; }

/* frames first_frame .. first_frame + frame_count - 1, each one vkCmdTraceRaysKHR with
 * SampleCount = samples_per_frame and TotalSamples = first_sample + f * samples_per_frame
 * (Renderer.cpp:1688-1700).  Returns the number of pixels whose restart loop did not end. */
PT_API int32_t glr_render(const glr_scene *h, const pt_render_params *p, uint32_t width, uint32_t height,
                          uint32_t first_sample, uint32_t frame_count, uint32_t samples_per_frame, float *accum,
                          int32_t threads)
{
    using namespace glslref;
    if (!h || !p || !accum)
        return PT_ERR_INVALID_ARGUMENT;
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1)
        nthreads = 1;
    std::atomic<uint32_t> nextRow { 0 };
    std::atomic<int32_t> spinning { 0 };
    auto worker = [&]() {
        bindScene(h->s, *p);
        tls_image = accum;
        gl_LaunchSizeEXT = uvec3(width, height, 1);
        for (;;)
        {
            const uint32_t y = nextRow.fetch_add(1);
            if (y >= height)
                break;
            for (uint32_t x = 0; x < width; x++)
                for (uint32_t f = 0; f < frame_count; f++)
                {
                    gl_LaunchIDEXT = uvec3(x, y, 0);
                    mainUniform.SampleCount = samples_per_frame;
                    mainUniform.TotalSamples = first_sample + f * samples_per_frame;
                    tls_trace_calls = 0;
                    std::memset(&payload, 0, sizeof(payload));
                    try
                    {
                        rgen::main_();
                    }
                    catch (const std::runtime_error &)
                    {
                        spinning++;
                    }
                }
        }
    };
    if (nthreads == 1)
        worker();
    else
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < nthreads; i++)
            pool.emplace_back(worker);
        for (auto &t : pool)
            t.join();
    }
    return spinning.load();
}

/* One frame of the debug pipeline (Debug/debugRaygen.rgen main() per pixel, dispatching to debugAnyhit.rahit /
 * debugClosestHit.rchit / debugMiss.rmiss and, for its shadow rays, occlusionAnyhit.rahit / occlusion.rmiss):
 * width * height RGBA floats. */
PT_API int32_t glr_debug_render(const glr_scene *h, const pt_render_params *p, const pt_debug_params *dbg, uint32_t width,
                                uint32_t height, float *out_rgba)
{
    using namespace glslref;
    if (!h || !p || !dbg || !out_rgba)
        return PT_ERR_INVALID_ARGUMENT;
    bindScene(h->s, *p);
    tls_image = out_rgba;
    gl_LaunchSizeEXT = uvec3(width, height, 1);
    s_RenderMode = dbg->render_mode;
    s_RaygenFlags = dbg->raygen_flags;
    s_HitGroupFlags = dbg->hit_group_flags;
    tls_debug_pipeline = true;
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++)
        {
            gl_LaunchIDEXT = uvec3(x, y, 0);
            tls_trace_calls = 0;
            std::memset(&dbg_payload, 0, sizeof(dbg_payload));
            dbg_rgen::main_();
        }
    tls_debug_pipeline = false;
    return PT_OK;
}

/* closestHit.rchit main() for `count` hits: rays = origin.xyz, direction.xyz (6 floats);
 * payloads = 36 floats in the layout of Shaders::Payload (ShaderRendererTypes.incl:101-118) */
PT_API int32_t glr_closest_hit(const glr_scene *h, const pt_render_params *p, uint32_t count, const pt_hit *hits,
                               const float *rays6, const float *payload_in, float *payload_out)
{
    using namespace glslref;
    if (!h || !p || !hits || !rays6 || !payload_in || !payload_out)
        return PT_ERR_INVALID_ARGUMENT;
    bindScene(h->s, *p);
    for (uint32_t i = 0; i < count; i++)
    {
        std::memcpy(&payload, payload_in + 36 * (size_t)i, sizeof(Payload));
        const float *r = rays6 + 6 * (size_t)i;
        gl_WorldRayOriginEXT = vec3(r[0], r[1], r[2]);
        gl_WorldRayDirectionEXT = vec3(r[3], r[4], r[5]);
        bindHit(h->s, hits[i].instance, hits[i].geometry, hits[i].primitive, hits[i].t, hits[i].u, hits[i].v);
        rchit::main_();
        std::memcpy(payload_out + 36 * (size_t)i, &payload, sizeof(Payload));
    }
    return PT_OK;
}

PT_API int32_t glr_test_shading(uint32_t mode, const float *in, float *out, uint32_t count)
{
    using namespace glslref;
    using namespace glslref::rchit;
    static const uint32_t kIn[PT_TEST_MODE_COUNT] = PT_TEST_INPUT_STRIDES, kOut[PT_TEST_MODE_COUNT] = PT_TEST_OUTPUT_STRIDES;
    if (mode >= PT_TEST_MODE_COUNT || !in || !out)
        return PT_ERR_INVALID_ARGUMENT;
    const uint32_t is = kIn[mode], os = kOut[mode];
    auto bits = [](float f) { return glm::floatBitsToUint(f); };
    for (uint32_t i = 0; i < count; i++)
    {
        const float *a = in + (size_t)i * is;
        float *o = out + (size_t)i * os;
        switch (mode)
        {
        case PT_TEST_GGX_DISTRIBUTION:
            o[0] = GGXDistribution(vec3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_LAMBDA:
            o[0] = Lambda(vec3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_GGX_SMITH:
            o[0] = GGXSmith(vec3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_DIELECTRIC_FRESNEL:
            o[0] = DielectricFresnel(a[0], a[1]);
            break;
        case PT_TEST_SCHLICK_FRESNEL:
            o[0] = SchlickFresnel(a[0]);
            break;
        case PT_TEST_EVALUATE_REFLECTION: {
            float pdf;
            const vec3 f = EvaluateReflection(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]), vec3(a[6], a[7], a[8]), a[9], pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_EVALUATE_REFRACTION: {
            float pdf;
            const vec3 f = EvaluateRefraction(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]), vec3(a[6], a[7], a[8]), a[9],
                                              a[10], pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_SAMPLE_GGX: {
            const vec3 hv = SampleGGX(vec2(a[0], a[1]), vec3(a[2], a[3], a[4]), a[5]);
            o[0] = hv.x, o[1] = hv.y, o[2] = hv.z;
            break;
        }
        case PT_TEST_SAMPLE_LOBE_PDFS: {
            MaterialSample m = {};
            m.Metalness = a[0];
            m.Transmission = a[1];
            const LobePdfs q = sampleLobePdfs(m, a[2]);
            o[0] = q.Diffuse, o[1] = q.Glossy, o[2] = q.Metallic, o[3] = q.Transmissive;
            break;
        }
        case PT_TEST_EVALUATE_BSDF: {
            const MaterialSample m = materialFromFloats(a);
            float pdf;
            const vec3 f = evaluateBSDF(m, vec3(a[17], a[18], a[19]), vec3(a[20], a[21], a[22]), pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_SAMPLE_BSDF: {
            const MaterialSample m = materialFromFloats(a);
            uint rng = bits(a[20]);
            const BSDFSample b = sampleBSDF(m, vec3(a[17], a[18], a[19]), rng);
            o[0] = b.Direction.x, o[1] = b.Direction.y, o[2] = b.Direction.z, o[3] = b.Pdf;
            o[4] = b.Color.x, o[5] = b.Color.y, o[6] = b.Color.z, o[7] = glm::uintBitsToFloat(rng);
            break;
        }
        case PT_TEST_RNG: {
            uint st = initRng(uvec2(bits(a[0]), bits(a[1])), uvec2(bits(a[2]), 0u), bits(a[3]));
            o[0] = glm::uintBitsToFloat(st);
            for (int k = 0; k < 4; k++)
            {
                const float f = rand(st);
                o[1 + k] = glm::uintBitsToFloat(st);
                o[5 + k] = f;
            }
            break;
        }
        case PT_TEST_PRIMARY_RAY: {
            const uvec2 pixel(bits(a[0]), bits(a[1])), res(bits(a[2]), bits(a[3]));
            Camera cam;
            std::memcpy(&cam.ViewInverse, a + 10, 64);
            std::memcpy(&cam.ProjInverse, a + 26, 64);
            Ray rx, ry, r;
            if (a[8] > 0)
                r = constructPrimaryRay(pixel, res, cam, vec2(a[4], a[5]), vec2(a[6], a[7]), a[8], a[9], rx, ry);
            else
                r = constructPrimaryRay(pixel, res, cam, vec2(a[4], a[5]), rx, ry);
            const Ray *rs[3] = { &r, &rx, &ry };
            for (int k = 0; k < 3; k++)
            {
                o[k * 6 + 0] = rs[k]->Origin.x, o[k * 6 + 1] = rs[k]->Origin.y, o[k * 6 + 2] = rs[k]->Origin.z;
                o[k * 6 + 3] = rs[k]->Direction.x, o[k * 6 + 4] = rs[k]->Direction.y, o[k * 6 + 5] = rs[k]->Direction.z;
            }
            break;
        }
        case PT_TEST_OFFSET_SELF_INTERSECTION: {
            const vec3 r = offsetRayOriginSelfIntersection(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_CONCENTRIC_DISK: {
            const vec2 r = sampleUniformDiskConcentric(vec2(a[0], a[1]));
            o[0] = r.x, o[1] = r.y;
            break;
        }
        case PT_TEST_TANGENT_SPACE: {
            const mat3 m = computeTangentSpace(vec3(a[0], a[1], a[2]));
            for (int k = 0; k < 3; k++)
                o[k * 3] = m[k].x, o[k * 3 + 1] = m[k].y, o[k * 3 + 2] = m[k].z;
            break;
        }
        case PT_TEST_DPN_DUV: {
            const Vertex v0 = vertexPUN(a), v1 = vertexPUN(a + 8), v2 = vertexPUN(a + 16);
            Vertex vertex = {};
            vertex.Tangent = vec3(a[24], a[25], a[26]);
            vertex.Bitangent = vec3(a[27], a[28], a[29]);
            vec3 dpdu, dpdv, dndu, dndv;
            computeDpnDuv(v0, v1, v2, vertex, dpdu, dpdv, dndu, dndv);
            const vec3 r[4] = { dpdu, dpdv, dndu, dndv };
            for (int k = 0; k < 4; k++)
                o[k * 3] = r[k].x, o[k * 3 + 1] = r[k].y, o[k * 3 + 2] = r[k].z;
            break;
        }
        case PT_TEST_DP_DXY: {
            vec3 dpdx, dpdy;
            computeDpDxy(vec3(a[0], a[1], a[2]), vec3(0.0f), vec3(0.0f), vec3(a[3], a[4], a[5]), vec3(a[6], a[7], a[8]),
                         vec3(a[9], a[10], a[11]), vec3(a[12], a[13], a[14]), vec3(a[15], a[16], a[17]), dpdx, dpdy);
            o[0] = dpdx.x, o[1] = dpdx.y, o[2] = dpdx.z, o[3] = dpdy.x, o[4] = dpdy.y, o[5] = dpdy.z;
            break;
        }
        case PT_TEST_DERIVATIVES: {
            const vec4 r = computeDerivatives(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]), vec3(a[6], a[7], a[8]),
                                              vec3(a[9], a[10], a[11]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z, o[3] = r.w;
            break;
        }
        case PT_TEST_REFLECTED_DIFFERENTIALS:
        case PT_TEST_REFRACTED_DIFFERENTIALS: {
            const bool refr = mode == PT_TEST_REFRACTED_DIFFERENTIALS;
            const float *b = a + 22 + (refr ? 1 : 0);
            vec3 rxO(b[0], b[1], b[2]), rxD(b[3], b[4], b[5]), ryO(b[6], b[7], b[8]), ryD(b[9], b[10], b[11]);
            const vec4 der(a[0], a[1], a[2], a[3]);
            const vec3 n(a[4], a[5], a[6]), pp(a[7], a[8], a[9]), viewDir(a[10], a[11], a[12]), outDir(a[13], a[14], a[15]);
            const vec3 dndu(a[16], a[17], a[18]), dndv(a[19], a[20], a[21]);
            if (refr)
                computeRefractedDifferentialRays(der, n, pp, viewDir, outDir, dndu, dndv, a[22], rxO, rxD, ryO, ryD);
            else
                computeReflectedDifferentialRays(der, n, pp, viewDir, outDir, dndu, dndv, rxO, rxD, ryO, ryD);
            const vec3 r[4] = { rxO, rxD, ryO, ryD };
            for (int k = 0; k < 4; k++)
                o[k * 3] = r[k].x, o[k * 3 + 1] = r[k].y, o[k * 3 + 2] = r[k].z;
            break;
        }
        case PT_TEST_SHADOW_TERMINATOR: {
            Vertex vertex = {}, v0 = {}, v1 = {}, v2 = {};
            vertex.Position = vec3(a[0], a[1], a[2]);
            v0.Position = vec3(a[3], a[4], a[5]), v0.Normal = vec3(a[6], a[7], a[8]);
            v1.Position = vec3(a[9], a[10], a[11]), v1.Normal = vec3(a[12], a[13], a[14]);
            v2.Position = vec3(a[15], a[16], a[17]), v2.Normal = vec3(a[18], a[19], a[20]);
            const vec3 r = offsetRayOriginShadowTerminator(vertex, v0, v1, v2, vec3(a[21], a[22], a[23]), a[24] != 0.0f);
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_SAMPLE_LIGHT: {
            /* u.xyz, position.xyz, directional colour + direction, one point light (colour, position,
             * attenuation c/l/q), light count (bits, 0 or 1) */
            DirectionalLight dl = {};
            dl.Color = vec3(a[6], a[7], a[8]), dl.Direction = vec3(a[9], a[10], a[11]);
            PointLight pl = {};
            pl.Color = vec3(a[12], a[13], a[14]), pl.Position = vec3(a[15], a[16], a[17]);
            pl.AttenuationConstant = a[18], pl.AttenuationLinear = a[19], pl.AttenuationQuadratic = a[20];
            u_LightCount = bits(a[21]) ? 1u : 0u;
            u_DirectionalLight = dl;
            u_Lights = &pl;
            float pdf;
            const LightSample l = sampleLight(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]), pdf);
            o[0] = l.Direction.x, o[1] = l.Direction.y, o[2] = l.Direction.z, o[3] = l.Distance;
            o[4] = l.Color.x, o[5] = l.Color.y, o[6] = l.Color.z, o[7] = l.Attenuation, o[8] = pdf;
            u_Lights = nullptr;
            break;
        }
        case PT_TEST_TRANSFORM_VERTEX: {
            /* position, normal, tangent, bitangent (12), mesh transform (12), object-to-world (12) */
            Vertex v = {};
            v.Position = vec3(a[0], a[1], a[2]), v.Normal = vec3(a[3], a[4], a[5]);
            v.Tangent = vec3(a[6], a[7], a[8]), v.Bitangent = vec3(a[9], a[10], a[11]);
            mat3x4 mesh;
            std::memcpy(&mesh, a + 12, 48);
            std::memcpy(&gl_ObjectToWorld3x4EXT, a + 24, 48);
            transforms = &mesh;
            const Vertex r = transform(v, 0);
            transforms = nullptr;
            o[0] = r.Position.x, o[1] = r.Position.y, o[2] = r.Position.z, o[3] = r.Normal.x, o[4] = r.Normal.y, o[5] = r.Normal.z;
            o[6] = r.Tangent.x, o[7] = r.Tangent.y, o[8] = r.Tangent.z, o[9] = r.Bitangent.x, o[10] = r.Bitangent.y,
            o[11] = r.Bitangent.z;
            break;
        }
        case PT_TEST_RECONSTRUCT_NORMAL: {
            const vec3 r = ReconstructNormalFromXY(vec3(a[0], a[1], a[2]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_HDR_TO_LDR: {
            const vec3 r = hdrToLdr(vec3(a[0], a[1], a[2]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        default:
            return PT_ERR_INVALID_ARGUMENT;
        }
    }
    return PT_OK;
}

} /* extern "C" */
