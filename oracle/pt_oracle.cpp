/*
 * pt_oracle.cpp — CPU oracle for the path-tracing hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * What it is: a literal, scalar restatement of the reference's GLSL ray-tracing stages,
 * function for function and in the same evaluation order (the RNG stream is threaded through
 * every stage), one pixel at a time like one raygen invocation.  Every
 * function cites the reference file:line it follows (paths relative to the reference checkout,
 * PT/ = Path-Tracing/).
 *
 * Parity pinning:
 *   - EVERYTHING the reference's shaders compute — common / ray / tracing / shading / bsdf / sampling / material .glsl and
 *     the main()s of raygen.rgen, closestHit.rchit, anyhit.rahit, occlusionAnyhit.rahit, miss.rmiss, occlusion.rmiss — is
 *     pinned BIT FOR BIT to the reference's own source: oracle/ref_overlay/build_glsl.sh compiles those files as C++
 *     (mechanical transform glsl2cpp.py, the reference's vendored glm) into oracle/_ref/libglsl_ref.so, and
 *     tests/test_oracle_vs_glsl.py compares every function (26 probe modes; the reference's PTT/TestData.h grids and 10^5
 *     random records each), closest-hit payloads and whole accumulation images with this file: zero differing records.
 *     Golden vectors of the compiled GLSL travel in tests/golden/glsl_vectors.npz.  Where GLSL leaves an algorithm open
 *     (inverse(mat4), the summation order of mat4 * vec4) this file follows glm, the reference's own host-side choice.
 *   - the COMPUTE stages — skinning.comp (skinVertex below) and the post-process chain (pt_oracle_post.cpp) — are pinned the
 *     same way: glsl2cpp.py --compute -> oracle/_ref/libglsl_comp_ref.so, tests/test_oracle_vs_glsl_compute.py, golden vectors
 *     in tests/golden/glsl_compute_vectors.npz; every float of every skinned vertex and of the composed / tone-mapped images.
 *   - the debug pipeline (debugPixel / pto_debug_render below) is pinned through the same library: Debug/debug*.r* compiled next
 *     to the path-tracing stages, every render mode and flag, bit for bit (tests/test_oracle_vs_glsl.py::test_debug_pipeline_*).
 *   - RNG: additionally pinned by the known answers derived from the integer spec (tests/test_oracle_rng.py).
 *   - struct layouts: pinned by PTT/PaddingTest.cpp literals (tests/test_layout.py).
 *   - ray/box, ray/triangle, BVH: PARITY UNPINNED — in the reference they run inside the Vulkan driver / RT hardware and
 *     there is no source or test (the compiled GLSL reaches them through a callback into this file).  The oracle's fp32
 *     watertight test (Woop, Benthin, Wald 2013) is validated against an fp64 brute force.
 *   - texture filtering (textureGrad, anisotropic), block decompression: PARITY UNPINNED — sampler hardware.  The oracle
 *     defines trilinear filtering per the GL 4.6 spec §8.14 and, behind pto_scene_set_sampler, the anisotropic example
 *     implementation of the Vulkan specification; closed forms in tests/test_oracle_textures.py.
 */
#include "pt_oracle.h"
#include "glsl_math.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

using namespace glsl;

namespace
{

/* ========================================================================= */
/* common.glsl                                                               */
/* ========================================================================= */

const float PI = 3.14159265359f; /* PT/Shaders/common.glsl:3 */

/* PT/Shaders/common.glsl:12-15 */
float maxComponent(vec3 rgb) { return max(rgb.x, max(rgb.y, rgb.z)); }
/* PT/Shaders/common.glsl:17-20 */
vec3 hdrToLdr(vec3 rgb) { return rgb / (1.0f + maxComponent(rgb)); }
/* PT/Shaders/common.glsl:22-25 */
vec3 computeBarycentricCoords(vec2 attribs) { return V3(1.0f - attribs.x - attribs.y, attribs.x, attribs.y); }

/* Shaders::Vertex, PT/Shaders/ShaderTypes.incl:40-47 */
struct Vertex
{
    vec3 Position;
    vec2 TexCoords;
    vec3 Normal;
    vec3 Tangent;
    vec3 Bitangent;
};

Vertex toVertex(const pt_vertex &v)
{
    Vertex r;
    r.Position = V3(v.position[0], v.position[1], v.position[2]);
    r.TexCoords = V2(v.texcoords[0], v.texcoords[1]);
    r.Normal = V3(v.normal[0], v.normal[1], v.normal[2]);
    r.Tangent = V3(v.tangent[0], v.tangent[1], v.tangent[2]);
    r.Bitangent = V3(v.bitangent[0], v.bitangent[1], v.bitangent[2]);
    return r;
}

/* PT/Shaders/common.glsl:102-122 */
vec2 interpolate(vec2 v1, vec2 v2, vec2 v3, vec3 b) { return v1 * b.x + v2 * b.y + v3 * b.z; }
vec3 interpolate(vec3 v1, vec3 v2, vec3 v3, vec3 b) { return v1 * b.x + v2 * b.y + v3 * b.z; }
Vertex interpolate(const Vertex &v1, const Vertex &v2, const Vertex &v3, vec3 b)
{
    Vertex v;
    v.Position = interpolate(v1.Position, v2.Position, v3.Position, b);
    v.TexCoords = interpolate(v1.TexCoords, v2.TexCoords, v3.TexCoords, b);
    v.Normal = interpolate(v1.Normal, v2.Normal, v3.Normal, b);
    v.Tangent = interpolate(v1.Tangent, v2.Tangent, v3.Tangent, b);
    v.Bitangent = interpolate(v1.Bitangent, v2.Bitangent, v3.Bitangent, b);
    return v;
}

/* PT/Shaders/common.glsl:133-141 */
uint32_t jenkinsHash(uint32_t x)
{
    x += x << 10;
    x ^= x >> 6;
    x += x << 3;
    x ^= x >> 11;
    x += x << 15;
    return x;
}
/* PT/Shaders/common.glsl:143-147 — dot(uvec2, uvec2) is an integer dot product (Q1) */
uint32_t initRng(uint32_t px, uint32_t py, uint32_t resX, uint32_t frame)
{
    uint32_t rngState = (px * 1u + py * resX) ^ jenkinsHash(frame);
    return jenkinsHash(rngState);
}
/* PT/Shaders/common.glsl:149-152 */
float uintToFloat(uint32_t x) { return uintBitsToFloat(0x3f800000u | (x >> 9)) - 1.0f; }
/* PT/Shaders/common.glsl:154-160 */
uint32_t xorshift(uint32_t &s)
{
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}
/* PT/Shaders/common.glsl:162-165 */
float rnd(uint32_t &s) { return uintToFloat(xorshift(s)); }

/* PT/Shaders/common.glsl:168-184 */
vec2 sampleUniformDiskConcentric(vec2 u)
{
    vec2 offset = 2.0f * u - 1.0f;
    if (offset.x == 0.0f && offset.y == 0.0f)
        return V2(0.0f, 0.0f);
    if (abs(offset.x) > abs(offset.y))
    {
        float theta = PI / 4 * (offset.y / offset.x);
        return offset.x * V2(std::cos(theta), std::sin(theta));
    }
    else
    {
        float theta = PI / 2 - PI / 4 * (offset.x / offset.y);
        return offset.y * V2(std::cos(theta), std::sin(theta));
    }
}
/* PT/Shaders/common.glsl:186-191 */
vec3 sampleCosineHemisphere(vec2 u)
{
    vec2 d = sampleUniformDiskConcentric(u);
    float z = std::sqrt(1 - d.x * d.x - d.y * d.y);
    return V3(d.x, d.y, z);
}
/* PT/Shaders/common.glsl:193-202 */
mat3 computeTangentSpace(vec3 normal)
{
    vec3 t1 = cross(normal, V3(1.0f, 0.0f, 0.0f));
    vec3 t2 = cross(normal, V3(0.0f, 1.0f, 0.0f));
    vec3 tangent = length(t1) > length(t2) ? t1 : t2;
    vec3 bitangent = cross(normal, tangent);
    return M3(normalize(tangent), normalize(bitangent), normal);
}

/* ========================================================================= */
/* shading.glsl                                                              */
/* ========================================================================= */

/* PT/Shaders/shading.glsl:3-14 — D is clamped to <= 1 by max(denom, 1) */
float GGXDistribution(vec3 H, float alpha)
{
    const float Hx2 = H.x * H.x;
    const float Hy2 = H.y * H.y;
    const float Hz2 = H.z * H.z;
    const float alpha2 = alpha * alpha;
    const float denom = PI * alpha2 * std::pow(Hx2 / alpha2 + Hy2 / alpha2 + Hz2, 2.0f);
    return 1.0f / max(denom, 1.0f);
}
/* PT/Shaders/shading.glsl:16-27 */
float Lambda(vec3 V, float alpha)
{
    const float Vx2 = V.x * V.x;
    const float Vy2 = V.y * V.y;
    const float Vz2 = abs(V.z) * abs(V.z);
    const float alpha2 = alpha * alpha;
    const float nom = std::sqrt(1.0f + (alpha2 * Vx2 + alpha2 * Vy2) / Vz2) - 1.0f;
    return nom / 2.0f;
}
/* PT/Shaders/shading.glsl:29-32 */
float GGXSmith(vec3 V, float alpha) { return 1.0f / (1.0f + Lambda(V, alpha)); }
/* PT/Shaders/shading.glsl:34-48 */
float DielectricFresnel(float VdotH, float eta)
{
    float cosThetaI = VdotH;
    float sinThetaT2 = eta * eta * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaT2 > 1.0f)
        return 1.0f;
    const float cosThetaT = std::sqrt(max(1.0f - sinThetaT2, 0.0f));
    const float rs = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
    const float rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return (rs * rs + rp * rp) / 2.0f;
}
/* PT/Shaders/shading.glsl:50-53 */
float SchlickFresnel(float VdotH) { return std::pow(clamp(1.0f - VdotH, 0.0f, 1.0f), 5.0f); }

/* PT/Shaders/shading.glsl:56-77 */
vec3 EvaluateReflection(vec3 V, vec3 L, vec3 F, float alpha, float &pdf)
{
    if (L.z < 0.00001f)
    {
        pdf = 0.0f;
        return V3(0.0f);
    }
    const vec3 H = normalize(V + L);
    const float VdotH = dot(V, H);
    const float D = GGXDistribution(H, alpha);
    const float Gv = GGXSmith(V, alpha);
    const float Gl = GGXSmith(L, alpha);
    const float G = Gv * Gl;
    const float Dv = (Gv * max(VdotH, 0.0f) * D) / V.z;
    pdf = Dv / (4.0f * VdotH);
    return (D * G * F) / (4.0f * V.z);
}

/* PT/Shaders/shading.glsl:80-108 */
vec3 EvaluateRefraction(vec3 V, vec3 L, vec3 F, float alpha, float eta, float &pdf)
{
    if (L.z > -0.00001f)
    {
        pdf = 0.0f;
        return V3(0.0f);
    }
    vec3 H = normalize(eta * V + L);
    if (H.z < 0.0f)
        H = -H;
    const float VdotH = dot(V, H);
    const float LdotH = dot(L, H);
    const float D = GGXDistribution(H, alpha);
    const float Gv = GGXSmith(V, alpha);
    const float Gl = GGXSmith(L, alpha);
    const float G = Gv * Gl;
    const float Dv = (Gv * abs(VdotH) * D) / V.z;
    const float denominator = LdotH + eta * VdotH;
    const float jacobian = (std::pow(eta, 2.0f) * abs(LdotH)) / std::pow(denominator, 2.0f);
    pdf = Dv * jacobian;
    return (abs(VdotH) / abs(V.z)) * (D * G * F) * jacobian;
}

/* PT/Shaders/shading.glsl:111-129 */
vec3 SampleGGX(vec2 u, vec3 V, float alpha)
{
    vec3 Vh = normalize(V3(alpha * V.x, alpha * V.y, abs(V.z)));
    const float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    const vec3 T1 = lensq > 0 ? V3(-Vh.y, Vh.x, 0) * inversesqrt(lensq) : V3(1, 0, 0);
    const vec3 T2 = cross(Vh, T1);
    const float r = std::sqrt(u.x);
    const float phi = 2.0f * PI * u.y;
    const float t1 = r * std::cos(phi);
    float t2 = r * std::sin(phi);
    const float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * std::sqrt(1.0f - t1 * t1) + s * t2;
    const vec3 Nh = t1 * T1 + t2 * T2 + std::sqrt(max(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(V3(alpha * Nh.x, alpha * Nh.y, max(0.0f, Nh.z)));
}

/* ========================================================================= */
/* bsdf.glsl                                                                 */
/* ========================================================================= */

/* Shaders::MaterialSample, PT/Shaders/ShaderRendererTypes.incl:129-140 */
struct MaterialSample
{
    vec3 EmissiveColor;
    vec3 Color;
    vec3 Normal;
    float Roughness;
    float Metalness;
    float Transmission;
    float Eta;
    vec3 AttenuationColor;
    float AttenuationDistance;
};

/* PT/Shaders/bsdf.glsl:4-9 */
struct BSDFSample
{
    vec3 Direction;
    float Pdf;
    vec3 Color;
};

/* PT/Shaders/bsdf.glsl:11-15 */
vec3 evaluateDiffuseBRDF(const MaterialSample &m, vec3, vec3 L, float &pdf)
{
    pdf = L.z * 1.0f / PI;
    return L.z * m.Color / PI;
}
/* PT/Shaders/bsdf.glsl:22-25 */
vec3 evaluateGlossyBSDF(const MaterialSample &m, vec3 V, vec3 L, float &pdf)
{
    return EvaluateReflection(V, L, V3(1.0f), m.Roughness * m.Roughness, pdf);
}
/* PT/Shaders/bsdf.glsl:32-37 */
vec3 evaluateMetallicBRDF(const MaterialSample &m, vec3 V, vec3 L, float &pdf)
{
    const vec3 H = normalize(V + L);
    const vec3 F0 = mix(m.Color, V3(1.0f), SchlickFresnel(dot(V, H)));
    return EvaluateReflection(V, L, F0, m.Roughness * m.Roughness, pdf);
}
/* PT/Shaders/bsdf.glsl:44-47 */
vec3 evaluateBTDF(const MaterialSample &m, vec3 V, vec3 L, float &pdf)
{
    return EvaluateRefraction(V, L, m.Color, m.Roughness * m.Roughness, m.Eta, pdf);
}

/* PT/Shaders/bsdf.glsl:54-70 */
struct LobePdfs
{
    float Diffuse, Glossy, Metallic, Transmissive;
};
LobePdfs sampleLobePdfs(const MaterialSample &m, float F)
{
    LobePdfs p;
    p.Diffuse = (1.0f - m.Metalness) * (1.0f - F) * (1.0f - m.Transmission);
    p.Glossy = (1.0f - m.Metalness) * F;
    p.Metallic = m.Metalness;
    p.Transmissive = (1.0f - m.Metalness) * (1.0f - F) * m.Transmission;
    return p;
}

/* PT/Shaders/bsdf.glsl:72-103 */
vec3 evaluateBSDF(const MaterialSample &m, vec3 V, vec3 L, float &outPdf)
{
    const bool isReflection = L.z > 0.0f;
    vec3 H = isReflection ? normalize(V + L) : normalize(m.Eta * V + L);
    const float FD = DielectricFresnel(abs(dot(V, H)), m.Eta);
    LobePdfs pdfs = sampleLobePdfs(m, FD);
    vec3 bsdf = V3(0.0f);
    outPdf = 0.0f;
    float pdf;
    if (isReflection)
    {
        bsdf += evaluateDiffuseBRDF(m, V, L, pdf) * pdfs.Diffuse;
        outPdf += pdf * pdfs.Diffuse;
        bsdf += evaluateGlossyBSDF(m, V, L, pdf) * pdfs.Glossy;
        outPdf += pdf * pdfs.Glossy;
        bsdf += evaluateMetallicBRDF(m, V, L, pdf) * pdfs.Metallic;
        outPdf += pdf * pdfs.Metallic;
    }
    else
    {
        bsdf += evaluateBTDF(m, V, L, pdf) * pdfs.Transmissive;
        outPdf += pdf * pdfs.Transmissive;
    }
    return bsdf;
}

/* PT/Shaders/bsdf.glsl:105-132.  vec2(rand, rand) is evaluated left to right (Appendix B). */
BSDFSample sampleBSDF(const MaterialSample &m, vec3 V, uint32_t &rngState)
{
    const float alpha = m.Roughness * m.Roughness;
    const float u0 = rnd(rngState);
    const float u1 = rnd(rngState);
    const vec3 H = SampleGGX(V2(u0, u1), V, alpha);
    const float FD = DielectricFresnel(abs(dot(V, H)), m.Eta);
    vec3 L;
    if (rnd(rngState) < m.Metalness)
        L = normalize(reflect(-V, H)); /* sampleMetallicBRDF :39-42 */
    else
    {
        if (rnd(rngState) < FD)
            L = normalize(reflect(-V, H)); /* sampleGlossyBSDF :27-30 */
        else
        {
            if (rnd(rngState) < m.Transmission)
                L = normalize(refract(-V, H, m.Eta)); /* sampleBTDF :49-52 */
            else
            {
                const float d0 = rnd(rngState);
                const float d1 = rnd(rngState);
                L = sampleCosineHemisphere(V2(d0, d1)); /* sampleDiffuseBRDF :17-20 */
            }
        }
    }
    BSDFSample ret;
    ret.Direction = L;
    ret.Color = evaluateBSDF(m, V, L, ret.Pdf);
    return ret;
}

/* ========================================================================= */
/* ray.glsl                                                                  */
/* ========================================================================= */

const float origin_const = 1.0f / 32.0f;   /* PT/Shaders/ray.glsl:3-5 */
const float float_scale = 1.0f / 65536.0f;
const float int_scale = 256.0f;

struct Ray
{
    vec3 Origin;
    float tmin;
    vec3 Direction;
    float tmax;
};

mat4 toMat4(const float *m)
{
    mat4 r;
    std::memcpy(&r, m, 64);
    return r;
}

/* PT/Shaders/ray.glsl:16-56 (thin lens) */
Ray constructPrimaryRayLens(vec2 pixel, vec2 resolution, const mat4 &ViewInverse, const mat4 &ProjInverse, vec2 u,
                            vec2 u2, float lensRadius, float focalDistance, Ray &rx, Ray &ry)
{
    const vec2 pixelCenter = pixel + u;
    const vec2 pixelCenterOffsetX = pixelCenter + V2(1.0f, 0.0f);
    const vec2 pixelCenterOffsetY = pixelCenter + V2(0.0f, 1.0f);
    const vec2 pLens = lensRadius * sampleUniformDiskConcentric(u2);
    const vec2 inUV = pixelCenter / resolution;
    vec2 d = inUV * 2.0f - 1.0f;
    const vec2 inUVOffsetX = pixelCenterOffsetX / resolution;
    vec2 dOffsetX = inUVOffsetX * 2.0f - 1.0f;
    const vec2 inUVOffsetY = pixelCenterOffsetY / resolution;
    vec2 dOffsetY = inUVOffsetY * 2.0f - 1.0f;

    vec3 originCameraSpace = V3(pLens.x, pLens.y, 0);
    vec3 origin = xyz(ViewInverse * V4(originCameraSpace, 1));

    vec3 target = xyz(ProjInverse * V4(d.x, d.y, 1, 1));
    float ft = focalDistance / target.z;
    vec3 pFocus = ft * target;
    vec3 direction = xyz(ViewInverse * V4(normalize(pFocus - originCameraSpace), 0));

    vec3 targetOffsetX = xyz(ProjInverse * V4(dOffsetX.x, dOffsetX.y, 1, 1));
    float ftOffsetX = focalDistance / targetOffsetX.z;
    vec3 pFocusOffsetX = ftOffsetX * targetOffsetX;
    vec3 directionOffsetX = xyz(ViewInverse * V4(normalize(pFocusOffsetX - originCameraSpace), 0));

    vec3 targetOffsetY = xyz(ProjInverse * V4(dOffsetY.x, dOffsetY.y, 1, 1));
    float ftOffsetY = focalDistance / targetOffsetY.z;
    vec3 pFocusOffsetY = ftOffsetY * targetOffsetY;
    vec3 directionOffsetY = xyz(ViewInverse * V4(normalize(pFocusOffsetY - originCameraSpace), 0));

    float tmin = 0.00001f;
    float tmax = 10000.0f;
    rx = Ray { origin, tmin, directionOffsetX, tmax };
    ry = Ray { origin, tmin, directionOffsetY, tmax };
    return Ray { origin, tmin, direction, tmax };
}

/* PT/Shaders/ray.glsl:58-85 (pinhole) */
Ray constructPrimaryRay(vec2 pixel, vec2 resolution, const mat4 &ViewInverse, const mat4 &ProjInverse, vec2 u, Ray &rx,
                        Ray &ry)
{
    const vec2 pixelCenter = pixel + u;
    const vec2 pixelCenterOffsetX = pixelCenter + V2(1.0f, 0.0f);
    const vec2 pixelCenterOffsetY = pixelCenter + V2(0.0f, 1.0f);
    const vec2 inUV = pixelCenter / resolution;
    vec2 d = inUV * 2.0f - 1.0f;
    const vec2 inUVOffsetX = pixelCenterOffsetX / resolution;
    vec2 dOffsetX = inUVOffsetX * 2.0f - 1.0f;
    const vec2 inUVOffsetY = pixelCenterOffsetY / resolution;
    vec2 dOffsetY = inUVOffsetY * 2.0f - 1.0f;

    vec3 origin = xyz(ViewInverse * V4(0, 0, 0, 1));
    vec3 target = xyz(ProjInverse * V4(d.x, d.y, 1, 1));
    vec3 direction = xyz(ViewInverse * V4(normalize(target), 0));
    vec3 targetOffsetX = xyz(ProjInverse * V4(dOffsetX.x, dOffsetX.y, 1, 1));
    vec3 targetOffsetY = xyz(ProjInverse * V4(dOffsetY.x, dOffsetY.y, 1, 1));
    vec3 directionOffsetX = xyz(ViewInverse * V4(normalize(targetOffsetX), 0));
    vec3 directionOffsetY = xyz(ViewInverse * V4(normalize(targetOffsetY), 0));

    float tmin = 0.00001f;
    float tmax = 10000.0f;
    rx = Ray { origin, tmin, directionOffsetX, tmax };
    ry = Ray { origin, tmin, directionOffsetY, tmax };
    return Ray { origin, tmin, direction, tmax };
}

/* PT/Shaders/ray.glsl:93-106 (Waechter & Binder, RT Gems ch. 6) */
vec3 offsetRayOriginSelfIntersection(vec3 origin, vec3 normal)
{
    const int32_t ofx = (int32_t)(int_scale * normal.x);
    const int32_t ofy = (int32_t)(int_scale * normal.y);
    const int32_t ofz = (int32_t)(int_scale * normal.z);
    vec3 p_i = V3(intBitsToFloat(floatBitsToInt(origin.x) + ((origin.x < 0) ? -ofx : ofx)),
                  intBitsToFloat(floatBitsToInt(origin.y) + ((origin.y < 0) ? -ofy : ofy)),
                  intBitsToFloat(floatBitsToInt(origin.z) + ((origin.z < 0) ? -ofz : ofz)));
    return V3((abs(origin.x) < origin_const) ? origin.x + float_scale * normal.x : p_i.x,
              (abs(origin.y) < origin_const) ? origin.y + float_scale * normal.y : p_i.y,
              (abs(origin.z) < origin_const) ? origin.z + float_scale * normal.z : p_i.z);
}

/* PT/Shaders/ray.glsl:109-131 (Hanika, RT Gems II ch. 4); v0..v2 are by-value copies in GLSL */
vec3 offsetRayOriginShadowTerminator(const Vertex &vertex, Vertex v0, Vertex v1, Vertex v2, vec3 bary, bool isRefracted)
{
    vec3 tmpu = vertex.Position - v0.Position;
    vec3 tmpv = vertex.Position - v1.Position;
    vec3 tmpw = vertex.Position - v2.Position;
    if (isRefracted)
    {
        v0.Normal *= -1.0f;
        v1.Normal *= -1.0f;
        v2.Normal *= -1.0f;
    }
    float dotu = min(0.0f, dot(tmpu, v0.Normal));
    float dotv = min(0.0f, dot(tmpv, v1.Normal));
    float dotw = min(0.0f, dot(tmpw, v2.Normal));
    tmpu -= dotu * v0.Normal;
    tmpv -= dotv * v1.Normal;
    tmpw -= dotw * v2.Normal;
    return vertex.Position + bary.x * tmpu + bary.y * tmpv + bary.z * tmpw;
}

/* ========================================================================= */
/* tracing.glsl                                                              */
/* ========================================================================= */

/* PT/Shaders/tracing.glsl:2-28 */
void computeDpnDuv(const Vertex &v0, const Vertex &v1, const Vertex &v2, const Vertex &vertex, vec3 &dpdu, vec3 &dpdv,
                   vec3 &dndu, vec3 &dndv)
{
    vec3 e1 = v1.Position - v0.Position;
    vec3 e2 = v2.Position - v0.Position;
    vec3 en1 = v1.Normal - v0.Normal;
    vec3 en2 = v2.Normal - v0.Normal;
    vec2 duv1 = v1.TexCoords - v0.TexCoords;
    vec2 duv2 = v2.TexCoords - v0.TexCoords;
    float det = duv1.x * duv2.y - duv2.x * duv1.y;
    if (abs(det) < 1e-8f)
    {
        dpdu = vertex.Tangent;
        dpdv = vertex.Bitangent;
        dndu = V3(0.0f);
        dndv = V3(0.0f);
    }
    else
    {
        float invDet = 1.0f / det;
        dpdu = (duv2.y * e1 - duv1.y * e2) * invDet;
        dpdv = (-duv2.x * e1 + duv1.x * e2) * invDet;
        dndu = (duv2.y * en1 - duv1.y * en2) * invDet;
        dndv = (-duv2.x * en1 + duv1.x * en2) * invDet;
    }
}

/* PT/Shaders/tracing.glsl:31-41 (origin/direction parameters are unused there too) */
void computeDpDxy(vec3 p, vec3 rxOrigin, vec3 rxDirection, vec3 ryOrigin, vec3 ryDirection, vec3 n, vec3 &dpdx,
                  vec3 &dpdy)
{
    float d = -dot(n, p);
    float tx = (-dot(n, rxOrigin) - d) / dot(n, rxDirection);
    vec3 px = rxOrigin + tx * rxDirection;
    float ty = (-dot(n, ryOrigin) - d) / dot(n, ryDirection);
    vec3 py = ryOrigin + ty * ryDirection;
    dpdx = px - p;
    dpdy = py - p;
}

/* PT/Shaders/tracing.glsl:44-50 — the only place the shaders ask for a fused multiply-add */
float differenceOfProducts(float a, float b, float c, float d)
{
    float cd = c * d;
    float dop = std::fma(a, b, -cd);
    float error = std::fma(-c, d, cd);
    return dop + error;
}

/* PT/Shaders/tracing.glsl:53-78 */
vec4 computeDerivatives(vec3 dpdx, vec3 dpdy, vec3 dpdu, vec3 dpdv)
{
    float ata00 = dot(dpdu, dpdu);
    float ata01 = dot(dpdu, dpdv);
    float ata11 = dot(dpdv, dpdv);
    float invDet = 1 / differenceOfProducts(ata00, ata11, ata01, ata01);
    invDet = isinf(invDet) ? 0.0f : invDet;
    float atb0x = dot(dpdu, dpdx);
    float atb1x = dot(dpdv, dpdx);
    float atb0y = dot(dpdu, dpdy);
    float atb1y = dot(dpdv, dpdy);
    float dudx = differenceOfProducts(ata11, atb0x, ata01, atb1x) * invDet;
    float dvdx = differenceOfProducts(ata00, atb1x, ata01, atb0x) * invDet;
    float dudy = differenceOfProducts(ata11, atb0y, ata01, atb1y) * invDet;
    float dvdy = differenceOfProducts(ata00, atb1y, ata01, atb0y) * invDet;
    dudx = isinf(dudx) ? 0.0f : clamp(dudx, -1e8f, 1e8f);
    dvdx = isinf(dvdx) ? 0.0f : clamp(dvdx, -1e8f, 1e8f);
    dudy = isinf(dudy) ? 0.0f : clamp(dudy, -1e8f, 1e8f);
    dvdy = isinf(dvdy) ? 0.0f : clamp(dvdy, -1e8f, 1e8f);
    return V4(dudx, dvdx, dudy, dvdy);
}

/* PT/Shaders/tracing.glsl:81-108 */
void computeReflectedDifferentialRays(vec4 derivatives, vec3 n, vec3 p, vec3 viewDir, vec3 reflectedDir, vec3 dndu,
                                      vec3 dndv, vec3 &rxOrigin, vec3 &rxDirection, vec3 &ryOrigin, vec3 &ryDirection)
{
    float dudx = derivatives.x, dvdx = derivatives.y, dudy = derivatives.z, dvdy = derivatives.w;
    vec3 dndx = dndu * dudx + dndv * dvdx;
    vec3 dndy = dndu * dudy + dndv * dvdy;
    float d = -dot(n, p);
    float tx = (-dot(n, rxOrigin) - d) / dot(n, rxDirection);
    vec3 px = rxOrigin + tx * rxDirection;
    float ty = (-dot(n, ryOrigin) - d) / dot(n, ryDirection);
    vec3 py = ryOrigin + ty * ryDirection;
    vec3 dwodx = -rxDirection - viewDir;
    vec3 dwody = -ryDirection - viewDir;
    rxOrigin = px;
    ryOrigin = py;
    float dwoDotn_dx = dot(dwodx, n) + dot(viewDir, dndx);
    float dwoDotn_dy = dot(dwody, n) + dot(viewDir, dndy);
    rxDirection = normalize(reflectedDir - dwodx + 2 * (dot(viewDir, n) * dndx + dwoDotn_dx * n));
    ryDirection = normalize(reflectedDir - dwody + 2 * (dot(viewDir, n) * dndy + dwoDotn_dy * n));
}

/* PT/Shaders/tracing.glsl:111-148 */
void computeRefractedDifferentialRays(vec4 derivatives, vec3 n, vec3 p, vec3 viewDir, vec3 refractedDir, vec3 dndu,
                                      vec3 dndv, float eta, vec3 &rxOrigin, vec3 &rxDirection, vec3 &ryOrigin,
                                      vec3 &ryDirection)
{
    float dudx = derivatives.x, dvdx = derivatives.y, dudy = derivatives.z, dvdy = derivatives.w;
    vec3 dndx = dndu * dudx + dndv * dvdx;
    vec3 dndy = dndu * dudy + dndv * dvdy;
    float d = -dot(n, p);
    float tx = (-dot(n, rxOrigin) - d) / dot(n, rxDirection);
    vec3 px = rxOrigin + tx * rxDirection;
    float ty = (-dot(n, ryOrigin) - d) / dot(n, ryDirection);
    vec3 py = ryOrigin + ty * ryDirection;
    vec3 dwodx = -rxDirection - viewDir;
    vec3 dwody = -ryDirection - viewDir;
    rxOrigin = px;
    ryOrigin = py;
    if (dot(viewDir, n) < 0.0f)
    {
        n = -n;
        dndx = -dndx;
        dndy = -dndy;
    }
    float dwoDotn_dx = dot(dwodx, n) + dot(viewDir, dndx);
    float dwoDotn_dy = dot(dwody, n) + dot(viewDir, dndy);
    float mu = dot(viewDir, n) / eta - abs(dot(refractedDir, n));
    float dmudx = dwoDotn_dx * (1.0f / eta + 1.0f / (eta * eta) * dot(viewDir, n) / dot(refractedDir, n));
    float dmudy = dwoDotn_dy * (1.0f / eta + 1.0f / (eta * eta) * dot(viewDir, n) / dot(refractedDir, n));
    rxDirection = normalize(refractedDir - eta * dwodx + (mu * dndx + dmudx * n));
    ryDirection = normalize(refractedDir - eta * dwody + (mu * dndy + dmudy * n));
}

/* ========================================================================= */
/* textures: sampler2D with the reference's sampler (PT/Renderer/Renderer.cpp:103-112):    */
/* linear mag/min, linear mip, repeat.  Anisotropy is hardware-defined and NOT modelled     */
/* (parity unpinned) — isotropic trilinear per GL 4.6 §8.14 is the definition here and in   */
/* the CUDA core.                                                                           */
/* ========================================================================= */

struct Luts
{
    float unorm[256];
    float srgb[256];
    float srgbMid[255]; /* midpoints between consecutive srgb[] entries, for encoding */
    Luts()
    {
        for (int i = 0; i < 256; i++)
        {
            const double c = i / 255.0;
            unorm[i] = (float)i / 255.0f;
            srgb[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
        for (int i = 0; i < 255; i++)
            srgbMid[i] = 0.5f * (srgb[i] + srgb[i + 1]);
    }
};
const Luts g_luts;

uint8_t encodeUnorm8(float v)
{
    if (!(v > 0.0f))
        return 0;
    if (v >= 1.0f)
        return 255;
    return (uint8_t)(v * 255.0f + 0.5f);
}
/* nearest sRGB code in LINEAR space: smallest i with v < mid[i] (binary search) */
uint8_t encodeSrgb8(float v)
{
    int lo = 0, hi = 255;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (v < g_luts.srgbMid[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return (uint8_t)lo;
}

struct TexLevel
{
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> rgba8;
    std::vector<float> rgbaf;
};

struct Texture
{
    bool isFloat = false;
    bool srgb = false;
    std::vector<TexLevel> levels;

    vec4 texel(uint32_t level, uint32_t x, uint32_t y) const
    {
        const TexLevel &l = levels[level];
        const size_t o = ((size_t)y * l.w + x) * 4;
        if (isFloat)
            return V4(l.rgbaf[o], l.rgbaf[o + 1], l.rgbaf[o + 2], l.rgbaf[o + 3]);
        const float *lut = srgb ? g_luts.srgb : g_luts.unorm;
        return V4(lut[l.rgba8[o]], lut[l.rgba8[o + 1]], lut[l.rgba8[o + 2]], g_luts.unorm[l.rgba8[o + 3]]);
    }
};

/* Mip chain: levels = floor(log2(max(w,h))) + 1 (PT/Renderer/Image.cpp:14-17), each level a linear
 * vkCmdBlitImage of the previous one (Image.cpp:264-305): destination texel centre mapped to the
 * source, bilinear with clamp-to-edge; filtering in linear space, stored back as 8-bit. */
/* one linear vkCmdBlitImage of level `k` of t into a dw x dh level */
TexLevel blitLinear(const Texture &t, uint32_t k, uint32_t dw, uint32_t dh)
{
    const TexLevel &src = t.levels[k];
    TexLevel dst;
    dst.w = dw;
    dst.h = dh;
    if (t.isFloat)
        dst.rgbaf.resize((size_t)dst.w * dst.h * 4);
    else
        dst.rgba8.resize((size_t)dst.w * dst.h * 4);
    const float sxScale = (float)src.w / (float)dst.w;
    const float syScale = (float)src.h / (float)dst.h;
    for (uint32_t y = 0; y < dst.h; y++)
        for (uint32_t x = 0; x < dst.w; x++)
        {
            const float sx = ((float)x + 0.5f) * sxScale - 0.5f;
            const float sy = ((float)y + 0.5f) * syScale - 0.5f;
            const float fx0 = std::floor(sx), fy0 = std::floor(sy);
            const float fx = sx - fx0, fy = sy - fy0;
            const int x0 = std::clamp((int)fx0, 0, (int)src.w - 1), x1 = std::clamp((int)fx0 + 1, 0, (int)src.w - 1);
            const int y0 = std::clamp((int)fy0, 0, (int)src.h - 1), y1 = std::clamp((int)fy0 + 1, 0, (int)src.h - 1);
            const vec4 t00 = t.texel(k, x0, y0), t10 = t.texel(k, x1, y0);
            const vec4 t01 = t.texel(k, x0, y1), t11 = t.texel(k, x1, y1);
            const vec4 top = t00 * (1.0f - fx) + t10 * fx;
            const vec4 bot = t01 * (1.0f - fx) + t11 * fx;
            const vec4 v = top * (1.0f - fy) + bot * fy;
            const size_t o = ((size_t)y * dst.w + x) * 4;
            if (t.isFloat)
            {
                dst.rgbaf[o] = v.x;
                dst.rgbaf[o + 1] = v.y;
                dst.rgbaf[o + 2] = v.z;
                dst.rgbaf[o + 3] = v.w;
            }
            else
            {
                dst.rgba8[o] = t.srgb ? encodeSrgb8(v.x) : encodeUnorm8(v.x);
                dst.rgba8[o + 1] = t.srgb ? encodeSrgb8(v.y) : encodeUnorm8(v.y);
                dst.rgba8[o + 2] = t.srgb ? encodeSrgb8(v.z) : encodeUnorm8(v.z);
                dst.rgba8[o + 3] = encodeUnorm8(v.w);
            }
        }
    return dst;
}

void buildMips(Texture &t)
{
    uint32_t w = t.levels[0].w, h = t.levels[0].h;
    uint32_t count = 1;
    for (uint32_t m = std::max(w, h); m > 1; m >>= 1)
        count++;
    for (uint32_t k = 1; k < count; k++)
        t.levels.push_back(blitLinear(t, k - 1, std::max(1u, w >> k), std::max(1u, h >> k)));
}

/* TextureUploader::DetermineMaxTextureSizes + UploadTexture (TextureUploader.cpp:408-415, 551-569): the largest extent a
 * texture may have — MaxTextureDataSize = 4096, halved while a full mip chain of that extent exceeds the per-texture
 * share of the budget (0 = ForceFullTextureSize) — and the integer factor a larger texture is scaled down by. */
struct TextureLimits
{
    uint32_t maxTextureSize = 4096;
    uint64_t budgetBytes = 0;
    uint32_t textureCount = 1;
    uint32_t maxExtent(uint32_t bytesPerTexel) const
    {
        uint32_t extent = maxTextureSize;
        if (budgetBytes == 0)
            return extent;
        const uint64_t perTexture = budgetBytes / std::max<uint64_t>(1, textureCount);
        auto chainBytes = [&](uint32_t e) {
            uint64_t n = 0;
            for (uint32_t m = e;; m >>= 1)
            {
                n += (uint64_t)std::max(1u, m) * std::max(1u, m) * bytesPerTexel;
                if (m <= 1)
                    break;
            }
            return n;
        };
        while (extent > 1 && chainBytes(extent) > perTexture)
            extent /= 2;
        return extent;
    }
};

/* Block-compressed formats (TextureFormat::BC1 / BC3 / BC5, PT/Scene.h:35-42; VK_FORMAT_BC1_RGBA /
 * BC3 / BC5 in PT/Renderer/TextureUploader.cpp:586-591), decoded per the format definition: 565
 * endpoints expanded by bit replication, interpolated palette entries = the exact rationals rounded
 * to the nearest 8-bit value.  PARITY UNPINNED: the reference leaves decoding to the sampler
 * hardware, whose interpolation precision is implementation-defined. */
void bcAlphaBlock(const uint8_t *b, uint8_t out[16])
{
    uint8_t pal[8];
    const uint32_t a0 = b[0], a1 = b[1];
    pal[0] = (uint8_t)a0;
    pal[1] = (uint8_t)a1;
    if (a0 > a1)
        for (uint32_t i = 1; i < 7; i++)
            pal[1 + i] = (uint8_t)(((7 - i) * a0 + i * a1 + 3) / 7);
    else
    {
        for (uint32_t i = 1; i < 5; i++)
            pal[1 + i] = (uint8_t)(((5 - i) * a0 + i * a1 + 2) / 5);
        pal[6] = 0;
        pal[7] = 255;
    }
    uint64_t bits = 0;
    for (int k = 0; k < 6; k++)
        bits |= (uint64_t)b[2 + k] << (8 * k);
    for (int t = 0; t < 16; t++)
        out[t] = pal[(bits >> (3 * t)) & 7u];
}

void bcColorBlock(const uint8_t *b, bool punchThrough, uint8_t out[16][4])
{
    const uint32_t c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
    uint32_t pal[4][4];
    auto expand = [](uint32_t c, uint32_t *o) {
        const uint32_t r = c >> 11, g = (c >> 5) & 63u, bl = c & 31u;
        o[0] = (r << 3) | (r >> 2);
        o[1] = (g << 2) | (g >> 4);
        o[2] = (bl << 3) | (bl >> 2);
        o[3] = 255;
    };
    expand(c0, pal[0]);
    expand(c1, pal[1]);
    for (int k = 0; k < 3; k++)
    {
        if (c0 > c1 || !punchThrough)
        {
            pal[2][k] = (2 * pal[0][k] + pal[1][k] + 1) / 3;
            pal[3][k] = (pal[0][k] + 2 * pal[1][k] + 1) / 3;
        }
        else
        {
            pal[2][k] = (pal[0][k] + pal[1][k] + 1) / 2;
            pal[3][k] = 0;
        }
    }
    pal[2][3] = 255;
    pal[3][3] = (c0 > c1 || !punchThrough) ? 255 : 0;
    const uint32_t idx = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
    for (int t = 0; t < 16; t++)
        for (int k = 0; k < 4; k++)
            out[t][k] = (uint8_t)pal[(idx >> (2 * t)) & 3u][k];
}

Texture makeBlockCompressedTexture(const pt_texture_desc &d)
{
    Texture t;
    t.isFloat = false;
    t.srgb = d.format == PT_TEXTURE_BC3 || (d.format == PT_TEXTURE_BC1 && d.srgb != 0);
    const uint8_t *src = (const uint8_t *)d.pixels;
    const uint32_t levels = d.levels ? d.levels : 1u;
    const size_t blockBytes = d.format == PT_TEXTURE_BC1 ? 8 : 16;
    for (uint32_t level = 0; level < levels; level++)
    {
        TexLevel l;
        l.w = std::max(1u, d.width >> level);
        l.h = std::max(1u, d.height >> level);
        l.rgba8.assign((size_t)l.w * l.h * 4, 0);
        const uint32_t bw = (l.w + 3) / 4, bh = (l.h + 3) / 4;
        for (uint32_t by = 0; by < bh; by++)
            for (uint32_t bx = 0; bx < bw; bx++, src += blockBytes)
            {
                uint8_t texel[16][4];
                if (d.format == PT_TEXTURE_BC5)
                {
                    uint8_t r[16], g[16];
                    bcAlphaBlock(src, r);
                    bcAlphaBlock(src + 8, g);
                    for (int k = 0; k < 16; k++)
                        texel[k][0] = r[k], texel[k][1] = g[k], texel[k][2] = 0, texel[k][3] = 255;
                }
                else if (d.format == PT_TEXTURE_BC3)
                {
                    uint8_t a[16];
                    bcAlphaBlock(src, a);
                    bcColorBlock(src + 8, false, texel);
                    for (int k = 0; k < 16; k++)
                        texel[k][3] = a[k];
                }
                else
                    bcColorBlock(src, true, texel);
                for (int k = 0; k < 16; k++)
                {
                    const uint32_t x = bx * 4 + (k & 3), y = by * 4 + (k >> 2);
                    if (x < l.w && y < l.h)
                        std::memcpy(&l.rgba8[((size_t)y * l.w + x) * 4], texel[k], 4);
                }
            }
        t.levels.push_back(std::move(l));
    }
    return t;
}

Texture makeTexture(const pt_texture_desc &d, const TextureLimits *limits = nullptr)
{
    if (d.format >= PT_TEXTURE_BC1)
        return makeBlockCompressedTexture(d);
    Texture t;
    t.isFloat = d.format == PT_TEXTURE_RGBAF32;
    t.srgb = !t.isFloat && d.srgb != 0;
    TexLevel l;
    l.w = d.width;
    l.h = d.height;
    const size_t n = (size_t)d.width * d.height * 4;
    if (t.isFloat)
        l.rgbaf.assign((const float *)d.pixels, (const float *)d.pixels + n);
    else
        l.rgba8.assign((const uint8_t *)d.pixels, (const uint8_t *)d.pixels + n);
    t.levels.push_back(std::move(l));
    if (limits && !t.isFloat)
    {
        const uint32_t e = limits->maxExtent(4);
        const uint32_t scale = std::max((d.width + e - 1) / e, (d.height + e - 1) / e);
        if (scale > 1) /* linear blit of the full-size staging image into the smaller level 0 */
        {
            TexLevel small = blitLinear(t, 0, std::max(d.width / scale, 1u), std::max(d.height / scale, 1u));
            t.levels[0] = std::move(small);
        }
    }
    buildMips(t);
    return t;
}

Texture makeDefaultTexture(uint32_t rgba, bool srgb)
{
    /* PT/Renderer/Renderer.cpp:127-173: 1x1 RGBA8 from a little-endian uint */
    pt_texture_desc d = {};
    d.width = d.height = 1;
    d.format = PT_TEXTURE_RGBA8;
    d.srgb = srgb;
    d.pixels = &rgba;
    return makeTexture(d);
}

struct TexelCounter
{
    uint64_t n = 0;
};

/* bilinear fetch of one level with repeat addressing, normalised coordinates */
vec4 sampleBilinear(const Texture &t, uint32_t level, vec2 uv, TexelCounter *tc)
{
    const TexLevel &l = t.levels[level];
    if (l.w == 1 && l.h == 1)
    {
        if (tc)
            tc->n += 1;
        return t.texel(level, 0, 0);
    }
    float u = std::isfinite(uv.x) ? uv.x : 0.0f;
    float v = std::isfinite(uv.y) ? uv.y : 0.0f;
    u = u - std::floor(u);
    v = v - std::floor(v);
    const float x = u * (float)l.w - 0.5f;
    const float y = v * (float)l.h - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int W = (int)l.w, H = (int)l.h;
    const int x0 = (((int)fx0 % W) + W) % W, x1 = (x0 + 1) % W;
    const int y0 = (((int)fy0 % H) + H) % H, y1 = (y0 + 1) % H;
    const vec4 t00 = t.texel(level, x0, y0), t10 = t.texel(level, x1, y0);
    const vec4 t01 = t.texel(level, x0, y1), t11 = t.texel(level, x1, y1);
    if (tc)
        tc->n += 4;
    const vec4 top = t00 * (1.0f - fx) + t10 * fx;
    const vec4 bot = t01 * (1.0f - fx) + t11 * fx;
    return top * (1.0f - fy) + bot * fy;
}

/* texture() outside a fragment shader has no implicit derivatives: level 0
 * (anyhit.rahit:51, occlusionAnyhit.rahit:50, miss.rmiss:27) */
vec4 textureLod0(const Texture &t, vec2 uv, TexelCounter *tc = nullptr) { return sampleBilinear(t, 0, uv, tc); }

/* One trilinear tap: levels floor(lambda) and floor(lambda) + 1 blended by the fraction. */
vec4 sampleTrilinear(const Texture &t, float lambda, uint32_t last, vec2 uv, TexelCounter *tc)
{
    if (!(lambda > 0.0f))
        return sampleBilinear(t, 0, uv, tc);
    if (lambda >= (float)last)
        return sampleBilinear(t, last, uv, tc);
    const float fl = std::floor(lambda);
    const uint32_t l0 = (uint32_t)fl;
    const float f = lambda - fl;
    const vec4 a = sampleBilinear(t, l0, uv, tc);
    if (!(f > 0.0f))
        return a;
    const vec4 b = sampleBilinear(t, l0 + 1, uv, tc);
    return a * (1.0f - f) + b * f;
}

/* textureGrad(sampler2D, P, dPdx, dPdy) with the reference's sampler (linear / linear-mip / repeat, anisotropy enabled at
 * the device maximum, PT/Renderer/Renderer.cpp:103-112).  Filtering is sampler HARDWARE in the reference (PARITY UNPINNED);
 * this is the example implementation the Vulkan specification gives ("Texel Anisotropic Filtering"):
 *     rho_x = |dPdx * size|, rho_y = |dPdy * size|, eta = min(rho_max / rho_min, maxAnisotropy), N = ceil(eta),
 *     lambda = log2(rho_max / eta),
 *     tau = 1/N * sum_{i=1..N} trilinear(P + (i / (N + 1) - 1/2) * dPd{major axis}, lambda)
 * maxAnisotropy = 1 is the isotropic trilinear filter of GL 4.6 §8.14 (N = 1, eta = 1, one tap at P). */
vec4 textureGrad(const Texture &t, vec2 uv, vec2 dPdx, vec2 dPdy, TexelCounter *tc, uint32_t maxAnisotropy = 1)
{
    const uint32_t last = (uint32_t)t.levels.size() - 1;
    if (last == 0)
        return sampleBilinear(t, 0, uv, tc);
    const float w = (float)t.levels[0].w, h = (float)t.levels[0].h;
    const float ax = dPdx.x * w, ay = dPdx.y * h;
    const float bx = dPdy.x * w, by = dPdy.y * h;
    const float rx2 = ax * ax + ay * ay, ry2 = bx * bx + by * by;
    const float rho2 = max(rx2, ry2);
    uint32_t N = 1;
    float eta = 1.0f;
    if (maxAnisotropy > 1 && rho2 > 0.0f) /* a point footprint (both derivatives zero) is one tap of level 0 */
    {
        const float rmin2 = min(rx2, ry2);
        /* rho_max / rho_min; a zero (or non-finite) minor axis asks for the maximum */
        const float ratio = std::sqrt(rho2) / std::sqrt(rmin2);
        eta = (ratio <= (float)maxAnisotropy) ? max(ratio, 1.0f) : (float)maxAnisotropy;
        N = (uint32_t)std::ceil(eta);
    }
    const float lambda = 0.5f * std::log2(rho2 / (eta * eta));
    /* at the 1 x 1 top level every tap reads the same texel: one tap */
    if (N == 1 || lambda >= (float)last)
        return sampleTrilinear(t, lambda, last, uv, tc);
    const vec2 major = (rx2 > ry2) ? dPdx : dPdy;
    vec4 acc = V4(0.0f, 0.0f, 0.0f, 0.0f);
    for (uint32_t i = 1; i <= N; i++)
    {
        const float o = (float)i / (float)(N + 1) - 0.5f;
        const vec4 tap = sampleTrilinear(t, lambda, last, V2(uv.x + o * major.x, uv.y + o * major.y), tc);
        acc = (i == 1) ? tap : acc + tap;
    }
    const float n = (float)N;
    return V4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);
}

/* ========================================================================= */
/* scene                                                                     */
/* ========================================================================= */

struct FlatTri
{
    vec3 p0, p1, p2;    /* world space */
    uint32_t instance;  /* gl_InstanceID */
    uint32_t geometry;  /* mesh index inside the model == gl_GeometryIndexEXT */
    uint32_t primitive; /* gl_PrimitiveID */
    uint32_t opaque;
};

struct BvhNode
{
    float bmin[3], bmax[3];
    uint32_t left;  /* internal: index of left child (right = left + 1); leaf: first triangle */
    uint32_t count; /* 0 => internal */
};

} // namespace

struct pto_scene
{
    std::vector<pt_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<mat3x4> transforms; /* GLSL `mat3x4 transforms[]`: column j = row j of the 3x4 matrix */
    std::vector<pt_geometry> geometries;
    std::vector<pt_mesh_record> meshRecords;
    std::vector<pt_model> models;
    std::vector<pt_instance> instances;
    std::vector<pt_material_mr> mr;
    std::vector<pt_material_sg> sg;
    std::vector<pt_material_phong> phong;
    std::vector<Texture> textures;
    uint32_t maxAnisotropy = 1; /* pto_scene_set_sampler: 16 = the reference's sampler state (Renderer.cpp:103-112) */
    std::vector<pt_point_light> pointLights;
    pt_directional_light directional;
    bool hasSky2D = false;
    Texture sky2D;
    bool hasSkyCube = false;
    Texture skyCube[6]; /* Vulkan layer order +X, -X, +Y, -Y, +Z, -Z */

    std::vector<FlatTri> tris;      /* flattening order: instance, mesh-in-model, primitive */
    std::vector<uint32_t> triOrder; /* BVH leaf order -> index into tris */
    std::vector<BvhNode> nodes;
};

namespace
{

/* gl_ObjectToWorld3x4EXT of an instance: mat3x4 whose column j is row j of the 3x4 matrix */
mat3x4 objectToWorld3x4(const pt_instance &inst)
{
    const float *m = inst.transform;
    return mat3x4 { { V4(m[0], m[1], m[2], m[3]), V4(m[4], m[5], m[6], m[7]), V4(m[8], m[9], m[10], m[11]) } };
}

/* PT/Shaders/sampling.glsl:5-15 */
Vertex transformVertex(const pto_scene &s, Vertex vertex, uint32_t transformIndex, const mat3x4 &objectToWorld)
{
    const mat3x4 transform = M4(s.transforms[transformIndex]) * objectToWorld;
    vertex.Position = V4(vertex.Position, 1.0f) * transform;
    vertex.Tangent = normalize(V4(vertex.Tangent, 0.0f) * transform);
    vertex.Bitangent = normalize(V4(vertex.Bitangent, 0.0f) * transform);
    vertex.Normal = normalize(xyz(V4(vertex.Normal, 0.0f) * transpose(inverse(M4(transform)))));
    return vertex;
}

/* PT/Shaders/common.glsl:27-46 — indices are relative to the geometry's first vertex */
Vertex getVertex(const pto_scene &s, const pt_geometry &g, uint32_t offset)
{
    const uint32_t index = s.indices[g.index_offset + offset];
    return toVertex(s.vertices[g.vertex_offset + index]);
}

/* skinning.comp:21-50 — one animated vertex through its (at most MaxBonesPerVertex = 4) bones.
 * boneTransforms[b] = GLSL mat3x4 (three vec4 columns). */
pt_vertex skinVertex(const pt_animated_vertex &a, const float *boneTransforms, uint32_t boneCount)
{
    const vec3 P = V3(a.position[0], a.position[1], a.position[2]);
    const vec3 N = V3(a.normal[0], a.normal[1], a.normal[2]);
    const vec3 T = V3(a.tangent[0], a.tangent[1], a.tangent[2]);
    const vec3 B = V3(a.bitangent[0], a.bitangent[1], a.bitangent[2]);
    vec3 position = V3(0.0f, 0.0f, 0.0f), normal = position, tangent = position, bitangent = position;
    float totalWeight = 0;
    for (int i = 0; i < 4 && totalWeight < 1.0f; i++)
    {
        const uint32_t boneIndex = std::min(a.bone_indices[i], boneCount - 1);
        const float boneWeight = a.bone_weights[i];
        const float *m = boneTransforms + 12 * (size_t)boneIndex;
        const mat3x4 transform = { { V4(m[0], m[1], m[2], m[3]), V4(m[4], m[5], m[6], m[7]), V4(m[8], m[9], m[10], m[11]) } };
        /* `boneWeight * vec4(Position, 1) * transform` groups from the left: the WEIGHTED point goes through the matrix
         * (skinning.comp:41; pinned by tests/test_oracle_vs_glsl_compute.py — weighting the transformed point instead
         * differs in the last bit of x / y for every second vertex) */
        position = position + V4(P.x * boneWeight, P.y * boneWeight, P.z * boneWeight, boneWeight) * transform;
        tangent = tangent + normalize(V4(T, 0.0f) * transform) * boneWeight;
        bitangent = bitangent + normalize(V4(B, 0.0f) * transform) * boneWeight;
        const vec4 n4 = V4(N, 0.0f) * transpose(inverse(M4(transform)));
        normal = normal + normalize(V3(n4.x, n4.y, n4.z)) * boneWeight;
        totalWeight += boneWeight;
    }
    pt_vertex v = {};
    v.position[0] = position.x, v.position[1] = position.y, v.position[2] = position.z;
    v.texcoords[0] = a.texcoords[0], v.texcoords[1] = a.texcoords[1];
    v.normal[0] = normal.x, v.normal[1] = normal.y, v.normal[2] = normal.z;
    v.tangent[0] = tangent.x, v.tangent[1] = tangent.y, v.tangent[2] = tangent.z;
    v.bitangent[0] = bitangent.x, v.bitangent[1] = bitangent.y, v.bitangent[2] = bitangent.z;
    return v;
}

/* ------------------------------------------------------------------------- */
/* BVH2, binned SAH (the reference has no BVH code: the Vulkan driver builds it) */
/* ------------------------------------------------------------------------- */

struct Aabb
{
    float mn[3], mx[3];
    void reset()
    {
        mn[0] = mn[1] = mn[2] = INFINITY;
        mx[0] = mx[1] = mx[2] = -INFINITY;
    }
    void grow(vec3 p)
    {
        const float v[3] = { p.x, p.y, p.z };
        for (int a = 0; a < 3; a++)
        {
            mn[a] = std::min(mn[a], v[a]);
            mx[a] = std::max(mx[a], v[a]);
        }
    }
    void grow(const Aabb &b)
    {
        for (int a = 0; a < 3; a++)
        {
            mn[a] = std::min(mn[a], b.mn[a]);
            mx[a] = std::max(mx[a], b.mx[a]);
        }
    }
    float area() const
    {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BvhBuilder
{
    pto_scene &s;
    std::vector<Aabb> triBox;
    std::vector<vec3> centroid;

    explicit BvhBuilder(pto_scene &scene) : s(scene) {}

    void build()
    {
        const size_t n = s.tris.size();
        s.triOrder.resize(n);
        triBox.resize(n);
        centroid.resize(n);
        for (size_t i = 0; i < n; i++)
        {
            s.triOrder[i] = (uint32_t)i;
            triBox[i].reset();
            triBox[i].grow(s.tris[i].p0);
            triBox[i].grow(s.tris[i].p1);
            triBox[i].grow(s.tris[i].p2);
            centroid[i] = V3(0.5f * (triBox[i].mn[0] + triBox[i].mx[0]), 0.5f * (triBox[i].mn[1] + triBox[i].mx[1]),
                             0.5f * (triBox[i].mn[2] + triBox[i].mx[2]));
        }
        s.nodes.clear();
        s.nodes.reserve(2 * n + 1);
        s.nodes.emplace_back();
        if (n == 0)
        {
            BvhNode &r = s.nodes[0];
            for (int a = 0; a < 3; a++)
            {
                r.bmin[a] = 0;
                r.bmax[a] = 0;
            }
            r.left = 0;
            r.count = 0;
            return;
        }
        subdivide(0, 0, (uint32_t)n);
    }

    void setBounds(uint32_t node, uint32_t first, uint32_t count)
    {
        Aabb b;
        b.reset();
        for (uint32_t i = first; i < first + count; i++)
            b.grow(triBox[s.triOrder[i]]);
        for (int a = 0; a < 3; a++)
        {
            s.nodes[node].bmin[a] = b.mn[a];
            s.nodes[node].bmax[a] = b.mx[a];
        }
    }

    void subdivide(uint32_t node, uint32_t first, uint32_t count)
    {
        setBounds(node, first, count);
        s.nodes[node].left = first;
        s.nodes[node].count = count;
        if (count <= 2)
            return;
        Aabb cb;
        cb.reset();
        for (uint32_t i = first; i < first + count; i++)
            cb.grow(centroid[s.triOrder[i]]);
        const int BINS = 16;
        float bestCost = INFINITY;
        int bestAxis = -1, bestSplit = -1;
        for (int axis = 0; axis < 3; axis++)
        {
            const float lo = cb.mn[axis], hi = cb.mx[axis];
            if (!(hi > lo))
                continue;
            Aabb binBox[BINS];
            uint32_t binCount[BINS] = {};
            for (auto &b : binBox)
                b.reset();
            const float scale = BINS / (hi - lo);
            for (uint32_t i = first; i < first + count; i++)
            {
                const uint32_t t = s.triOrder[i];
                int b = std::min(BINS - 1, (int)((centroid[t][axis] - lo) * scale));
                binCount[b]++;
                binBox[b].grow(triBox[t]);
            }
            float leftArea[BINS - 1], rightArea[BINS - 1];
            uint32_t leftN[BINS - 1], rightN[BINS - 1];
            Aabb l, r;
            l.reset();
            r.reset();
            uint32_t ln = 0, rn = 0;
            for (int i = 0; i < BINS - 1; i++)
            {
                ln += binCount[i];
                l.grow(binBox[i]);
                leftN[i] = ln;
                leftArea[i] = l.area();
                rn += binCount[BINS - 1 - i];
                r.grow(binBox[BINS - 1 - i]);
                rightN[BINS - 2 - i] = rn;
                rightArea[BINS - 2 - i] = r.area();
            }
            for (int i = 0; i < BINS - 1; i++)
            {
                if (leftN[i] == 0 || rightN[i] == 0)
                    continue;
                const float cost = leftN[i] * leftArea[i] + rightN[i] * rightArea[i];
                if (cost < bestCost)
                {
                    bestCost = cost;
                    bestAxis = axis;
                    bestSplit = i;
                }
            }
        }
        uint32_t mid;
        if (bestAxis < 0)
        {
            if (count <= 4)
                return;
            mid = first + count / 2; /* all centroids coincide: median split */
        }
        else
        {
            Aabb nb;
            for (int a = 0; a < 3; a++)
            {
                nb.mn[a] = s.nodes[node].bmin[a];
                nb.mx[a] = s.nodes[node].bmax[a];
            }
            const float leafCost = count * nb.area();
            if (count <= 4 && bestCost >= leafCost)
                return;
            const float lo = cb.mn[bestAxis], hi = cb.mx[bestAxis];
            const float scale = BINS / (hi - lo);
            auto it = std::partition(s.triOrder.begin() + first, s.triOrder.begin() + first + count, [&](uint32_t t) {
                int b = std::min(BINS - 1, (int)((centroid[t][bestAxis] - lo) * scale));
                return b <= bestSplit;
            });
            mid = (uint32_t)(it - s.triOrder.begin());
            if (mid == first || mid == first + count)
                mid = first + count / 2;
        }
        const uint32_t left = (uint32_t)s.nodes.size();
        s.nodes.emplace_back();
        s.nodes.emplace_back();
        s.nodes[node].left = left;
        s.nodes[node].count = 0;
        subdivide(left, first, mid - first);
        subdivide(left + 1, mid, first + count - mid);
    }
};

/* ------------------------------------------------------------------------- */
/* watertight ray/triangle (Woop, Benthin, Wald, JCGT 2013), fp32 with the    */
/* paper's fp64 fallback for zero edge functions                              */
/* ------------------------------------------------------------------------- */

struct RayPrep
{
    int kx, ky, kz;
    float Sx, Sy, Sz;
    vec3 org;
    float invDir[3];
    float orgA[3];
};

RayPrep prepareRay(vec3 org, vec3 dir)
{
    RayPrep r;
    const float ad[3] = { std::fabs(dir.x), std::fabs(dir.y), std::fabs(dir.z) };
    r.kz = (ad[0] > ad[1]) ? ((ad[0] > ad[2]) ? 0 : 2) : ((ad[1] > ad[2]) ? 1 : 2);
    r.kx = (r.kz + 1) % 3;
    r.ky = (r.kx + 1) % 3;
    if (dir[r.kz] < 0.0f)
        std::swap(r.kx, r.ky);
    r.Sx = dir[r.kx] / dir[r.kz];
    r.Sy = dir[r.ky] / dir[r.kz];
    r.Sz = 1.0f / dir[r.kz];
    r.org = org;
    r.invDir[0] = 1.0f / dir.x;
    r.invDir[1] = 1.0f / dir.y;
    r.invDir[2] = 1.0f / dir.z;
    r.orgA[0] = org.x;
    r.orgA[1] = org.y;
    r.orgA[2] = org.z;
    return r;
}

/* returns true and (t, b1, b2) when the ray hits; no culling; t range is checked by the caller */
bool intersectTriangle(const RayPrep &r, const FlatTri &tri, float &t, float &b1, float &b2)
{
    const vec3 A = tri.p0 - r.org, B = tri.p1 - r.org, C = tri.p2 - r.org;
    const float Ax = A[r.kx] - r.Sx * A[r.kz], Ay = A[r.ky] - r.Sy * A[r.kz];
    const float Bx = B[r.kx] - r.Sx * B[r.kz], By = B[r.ky] - r.Sy * B[r.kz];
    const float Cx = C[r.kx] - r.Sx * C[r.kz], Cy = C[r.ky] - r.Sy * C[r.kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    const float det = U + V + W;
    if (det == 0.0f)
        return false;
    const float Az = r.Sz * A[r.kz], Bz = r.Sz * B[r.kz], Cz = r.Sz * C[r.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float rcpDet = 1.0f / det;
    t = T * rcpDet;
    b1 = V * rcpDet;
    b2 = W * rcpDet;
    return true;
}

/* slab test; returns entry distance or INFINITY.  Conservative: uses <= so that ties are kept. */
float intersectBox(const RayPrep &r, const float *bmin, const float *bmax, float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++)
    {
        const float ta = (bmin[a] - r.orgA[a]) * r.invDir[a];
        const float tb = (bmax[a] - r.orgA[a]) * r.invDir[a];
        /* fmin/fmax drop NaNs (0 * inf when the origin lies on a slab of a flat box) */
        t0 = std::fmax(t0, std::fmin(ta, tb));
        t1 = std::fmin(t1, std::fmax(ta, tb));
    }
    /* widen by 2 ulp-ish so that rounding in the slab test can never cull a true hit */
    return (t0 <= t1 * 1.0000004f) ? t0 : INFINITY;
}

struct Counters
{
    uint64_t rays_closest = 0, rays_shadow = 0, samples = 0, hits = 0;
    uint64_t box_c = 0, tri_c = 0, box_s = 0, tri_s = 0, alpha_c = 0, alpha_s = 0, texels = 0, restarts = 0;
    void add(const Counters &o)
    {
        rays_closest += o.rays_closest;
        rays_shadow += o.rays_shadow;
        samples += o.samples;
        hits += o.hits;
        box_c += o.box_c;
        tri_c += o.tri_c;
        box_s += o.box_s;
        tri_s += o.tri_s;
        alpha_c += o.alpha_c;
        alpha_s += o.alpha_s;
        texels += o.texels;
        restarts += o.restarts;
    }
};

/* ------------------------------------------------------------------------- */
/* material.glsl helpers shared by the any-hit stages                          */
/* ------------------------------------------------------------------------- */

/* PT/Shaders/ShaderTypes.incl:162-167 */
uint32_t unpackMaterialId(uint32_t materialId, uint32_t &materialType)
{
    materialType = materialId & 0xffu;
    return materialId >> 8;
}
/* PT/Shaders/material.glsl:25-38 */
uint32_t getColorTextureIdx(const pto_scene &s, uint32_t materialIndex, uint32_t materialType)
{
    switch (materialType)
    {
    case PT_MATERIAL_METALLIC_ROUGHNESS:
        return s.mr[materialIndex].color_idx;
    case PT_MATERIAL_SPECULAR_GLOSSINESS:
        return s.sg[materialIndex].color_idx;
    case PT_MATERIAL_PHONG:
        return s.phong[materialIndex].color_idx;
    default:
        return 0;
    }
}
/* PT/Shaders/material.glsl:40-53 */
vec4 getColorFactor(const pto_scene &s, uint32_t materialIndex, uint32_t materialType)
{
    const float *c;
    switch (materialType)
    {
    case PT_MATERIAL_METALLIC_ROUGHNESS:
        c = s.mr[materialIndex].color;
        break;
    case PT_MATERIAL_SPECULAR_GLOSSINESS:
        c = s.sg[materialIndex].color;
        break;
    case PT_MATERIAL_PHONG:
        c = s.phong[materialIndex].color;
        break;
    default:
        return V4(1.0f, 0.0f, 0.0f, 1.0f);
    }
    return V4(c[0], c[1], c[2], c[3]);
}

const pt_mesh_record &meshRecordOf(const pto_scene &s, const FlatTri &t)
{
    /* instanceShaderBindingTableRecordOffset = MeshOffset * 2, stride 2, + geometry index
     * (PT/Renderer/AccelerationStructure.cpp:268-275, raygen.rgen:31,68) */
    return s.meshRecords[s.models[s.instances[t.instance].model_index].mesh_offset + t.geometry];
}

/* colour texture x colour factor at the candidate hit — the common part of anyhit.rahit:36-51 and
 * occlusionAnyhit.rahit:35-50 (full vertex interpolation, only TexCoords used) */
vec4 anyHitColor(const pto_scene &s, const FlatTri &t, float b1, float b2)
{
    const vec3 bary = computeBarycentricCoords(V2(b1, b2));
    const pt_mesh_record &rec = meshRecordOf(s, t);
    const pt_geometry &g = s.geometries[rec.geometry_index];
    const Vertex vertex = interpolate(getVertex(s, g, t.primitive * 3), getVertex(s, g, t.primitive * 3 + 1),
                                      getVertex(s, g, t.primitive * 3 + 2), bary);
    uint32_t materialType;
    const uint32_t materialIndex = unpackMaterialId(rec.material_id, materialType);
    const uint32_t colorTextureIdx = getColorTextureIdx(s, materialIndex, materialType);
    const vec4 colorFactor = getColorFactor(s, materialIndex, materialType);
    return textureLod0(s.textures[colorTextureIdx], vertex.TexCoords) * colorFactor;
}

/* ------------------------------------------------------------------------- */
/* traceRayEXT                                                                */
/* ------------------------------------------------------------------------- */

struct HitInfo
{
    uint32_t tri = PT_NO_HIT; /* index into scene.tris */
    float t = 0, b1 = 0, b2 = 0;
    /* decal record written by anyhit.rahit:54-62 (payload aliases, ShaderRendererTypes.incl:112-114) */
    float decalDist = -1.0f;
    vec3 decalColor = { 0, 0, 0 };
    float decalAlpha = 0.0f;
};

/* A ray with a NaN/Inf component or a zero direction can never satisfy a triangle test (everything
 * evaluates to NaN) but would visit every node, NaN defeating the slab test.  refract() == 0 on
 * total internal reflection followed by normalize() produces exactly such rays (Q12).  Vulkan
 * leaves non-finite rays undefined; oracle and core both define them as a miss. */
bool rayCanHit(vec3 org, vec3 dir)
{
    const float sum = org.x + org.y + org.z + dir.x + dir.y + dir.z;
    return std::isfinite(sum) && !(dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f);
}

/* closest hit with the primary any-hit shader.  Ties in t are broken towards the smaller flattened
 * triangle index so that the result does not depend on the BVH (the hardware's choice is
 * implementation-defined). */
/* External any-hit stage (pto_trace_anyhit): the candidate goes to the hook instead of the inline
 * restatement of anyhit.rahit / occlusionAnyhit.rahit; returns 1 = accept, 0 = ignoreIntersectionEXT. */
struct AnyHitHook
{
    pto_anyhit_fn fn = nullptr;
    void *ctx = nullptr;
};

/* gl_RayFlagsCullBackFacingTrianglesEXT (Debug/debugRaygen.rgen:32-35).  Vulkan decides the facing in OBJECT space (the
 * BLAS's space: after the mesh transform, before the instance transform), front = the vertices appear clockwise from
 * the ray origin = dot((v1 - v0) x (v2 - v0), d) < 0; the instance flags are empty (AccelerationStructure.cpp:273), so
 * nothing flips it.  With world-space data: an instance transform of negative determinant reverses the sign. */
bool isBackFacing(const pto_scene &s, const FlatTri &tri, vec3 dir)
{
    const vec3 n = cross(tri.p1 - tri.p0, tri.p2 - tri.p0);
    float facing = dot(n, dir);
    const float *I = s.instances[tri.instance].transform;
    const double det = (double)I[0] * ((double)I[5] * I[10] - (double)I[6] * I[9]) - (double)I[1] * ((double)I[4] * I[10] - (double)I[6] * I[8]) +
                       (double)I[2] * ((double)I[4] * I[9] - (double)I[5] * I[8]);
    if (det < 0.0)
        facing = -facing;
    return !(facing < 0.0f);
}

HitInfo traceClosest(const pto_scene &s, vec3 org, vec3 dir, float tmin, float tmax, Counters &c, bool forceOpaque = false,
                     const AnyHitHook *hook = nullptr, bool cullBackFaces = false)
{
    HitInfo hit;
    c.rays_closest++;
    if (s.tris.empty() || !rayCanHit(org, dir))
        return hit;
    const RayPrep r = prepareRay(org, dir);
    float best = tmax;
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    c.box_c++;
    if (intersectBox(r, s.nodes[0].bmin, s.nodes[0].bmax, tmin, best) == INFINITY)
        return hit;
    while (sp > 0)
    {
        const BvhNode &n = s.nodes[stack[--sp]];
        if (n.count == 0)
        {
            const BvhNode &l = s.nodes[n.left], &rr = s.nodes[n.left + 1];
            c.box_c += 2;
            const float tl = intersectBox(r, l.bmin, l.bmax, tmin, best);
            const float tr = intersectBox(r, rr.bmin, rr.bmax, tmin, best);
            if (tl != INFINITY && tr != INFINITY)
            {
                if (tl <= tr)
                {
                    stack[sp++] = n.left + 1;
                    stack[sp++] = n.left;
                }
                else
                {
                    stack[sp++] = n.left;
                    stack[sp++] = n.left + 1;
                }
            }
            else if (tl != INFINITY)
                stack[sp++] = n.left;
            else if (tr != INFINITY)
                stack[sp++] = n.left + 1;
            continue;
        }
        /* NOTE: a node pushed earlier may now lie beyond `best`; re-testing it costs more than it saves here */
        for (uint32_t i = n.left; i < n.left + n.count; i++)
        {
            const uint32_t ti = s.triOrder[i];
            const FlatTri &tri = s.tris[ti];
            float t, b1, b2;
            c.tri_c++;
            if (!intersectTriangle(r, tri, t, b1, b2))
                continue;
            if (!(t > tmin))
                continue;
            if (!(t < best || (t == best && hit.tri != PT_NO_HIT && ti < hit.tri)))
                continue;
            if (cullBackFaces && isBackFacing(s, tri, dir))
                continue;
            if (!tri.opaque && !forceOpaque) /* gl_RayFlagsOpaqueEXT skips the any-hit stage */
            {
                /* anyhit.rahit:36-65 */
                c.alpha_c++;
                if (hook)
                {
                    if (!hook->fn(hook->ctx, tri.instance, tri.geometry, tri.primitive, t, b1, b2))
                        continue;
                    best = t;
                    hit.tri = ti;
                    hit.t = t;
                    hit.b1 = b1;
                    hit.b2 = b2;
                    continue;
                }
                const vec4 color = anyHitColor(s, tri, b1, b2);
                if (color.w < 0.5f)
                {
                    if (hit.decalDist == -1.0f || t < hit.decalDist)
                    {
                        hit.decalColor = xyz(color);
                        hit.decalAlpha = color.w;
                        hit.decalDist = t;
                    }
                    continue; /* ignoreIntersectionEXT */
                }
            }
            best = t;
            hit.tri = ti;
            hit.t = t;
            hit.b1 = b1;
            hit.b2 = b2;
        }
    }
    if (hit.tri != PT_NO_HIT)
        c.hits++;
    return hit;
}

/* raygen.rgen:22-34 checkOccluded's traceRayEXT: TerminateOnFirstHit, occlusionAnyhit.rahit, occlusion.rmiss */
bool traceOccluded(const pto_scene &s, vec3 org, vec3 dir, float tmin, float tmax, Counters &c,
                   const AnyHitHook *hook = nullptr, HitInfo *outHit = nullptr)
{
    c.rays_shadow++;
    if (s.tris.empty() || !rayCanHit(org, dir))
        return false;
    const RayPrep r = prepareRay(org, dir);
    uint32_t stack[128];
    int sp = 0;
    c.box_s++;
    if (intersectBox(r, s.nodes[0].bmin, s.nodes[0].bmax, tmin, tmax) == INFINITY)
        return false;
    stack[sp++] = 0;
    while (sp > 0)
    {
        const BvhNode &n = s.nodes[stack[--sp]];
        if (n.count == 0)
        {
            const BvhNode &l = s.nodes[n.left], &rr = s.nodes[n.left + 1];
            c.box_s += 2;
            if (intersectBox(r, l.bmin, l.bmax, tmin, tmax) != INFINITY)
                stack[sp++] = n.left;
            if (intersectBox(r, rr.bmin, rr.bmax, tmin, tmax) != INFINITY)
                stack[sp++] = n.left + 1;
            continue;
        }
        for (uint32_t i = n.left; i < n.left + n.count; i++)
        {
            const FlatTri &tri = s.tris[s.triOrder[i]];
            float t, b1, b2;
            c.tri_s++;
            if (!intersectTriangle(r, tri, t, b1, b2))
                continue;
            if (!(t > tmin && t < tmax))
                continue;
            if (!tri.opaque)
            {
                /* occlusionAnyhit.rahit:35-54 */
                c.alpha_s++;
                if (hook)
                {
                    if (!hook->fn(hook->ctx, tri.instance, tri.geometry, tri.primitive, t, b1, b2))
                        continue;
                }
                else
                {
                    const float alpha = anyHitColor(s, tri, b1, b2).w;
                    if (alpha < 1.0f)
                        continue;
                }
            }
            if (outHit)
            {
                outHit->tri = s.triOrder[i];
                outHit->t = t;
                outHit->b1 = b1;
                outHit->b2 = b2;
            }
            return true;
        }
    }
    return false;
}

/* ========================================================================= */
/* material.glsl                                                             */
/* ========================================================================= */

/* PT/Shaders/material.glsl:55-60 */
vec3 ReconstructNormalFromXY(vec3 normal)
{
    normal = 2.0f * normal - 1.0f;
    return V3(normal.x, normal.y, std::sqrt(max(1 - normal.x * normal.x - normal.y * normal.y, 0.0f)));
}

struct TexCtx
{
    const pto_scene &s;
    vec2 uv, dpdx, dpdy;
    TexelCounter tc;
    vec4 grad(uint32_t idx) { return textureGrad(s.textures[idx], uv, dpdx, dpdy, &tc, s.maxAnisotropy); }
};

/* PT/Shaders/material.glsl:62-84 */
MaterialSample sampleMaterialMR(const pt_material_mr &m, TexCtx &tx, bool isHitFromInside, uint32_t flags = 0)
{
    MaterialSample ret;
    /* sampleValue(flags, HitGroupFlagsDisable*Texture, idx, Default*TextureIndex), material.glsl:4-23, 69-70 */
    const uint32_t colorIdx = (flags & PT_DEBUG_HIT_DISABLE_COLOR_TEXTURE) ? 0u : m.color_idx;
    const uint32_t normalIdx = (flags & PT_DEBUG_HIT_DISABLE_NORMAL_TEXTURE) ? 1u : m.normal_idx;
    const vec3 EmissiveColor = V3(m.emissive_color[0], m.emissive_color[1], m.emissive_color[2]);
    ret.EmissiveColor = (xyz(tx.grad(m.emissive_idx)) + EmissiveColor) * m.emissive_intensity;
    ret.Color = xyz(tx.grad(colorIdx)) * V3(m.color[0], m.color[1], m.color[2]);
    ret.Normal = ReconstructNormalFromXY(xyz(tx.grad(normalIdx)));
    ret.Roughness = tx.grad(m.roughness_idx).y * m.roughness;
    ret.Metalness = tx.grad(m.metallic_idx).z * m.metalness;
    ret.Transmission = m.transmission;
    ret.AttenuationColor = V3(m.attenuation_color[0], m.attenuation_color[1], m.attenuation_color[2]);
    ret.AttenuationDistance = m.attenuation_distance;
    ret.Eta = isHitFromInside ? m.ior : (1.0f / m.ior);
    return ret;
}

/* PT/Shaders/material.glsl:86-113 (specular-glossiness) and :115-142 (Phong: same code with
 * Shininess / ShininessIdx in the Glossiness slots — identical struct layout) */
MaterialSample sampleMaterialSG(const pt_material_sg &m, TexCtx &tx, bool isHitFromInside, uint32_t flags = 0)
{
    MaterialSample ret;
    const uint32_t colorIdx = (flags & PT_DEBUG_HIT_DISABLE_COLOR_TEXTURE) ? 0u : m.color_idx;
    const uint32_t normalIdx = (flags & PT_DEBUG_HIT_DISABLE_NORMAL_TEXTURE) ? 1u : m.normal_idx;
    const vec3 EmissiveColor = V3(m.emissive_color[0], m.emissive_color[1], m.emissive_color[2]);
    ret.EmissiveColor = (xyz(tx.grad(m.emissive_idx)) + EmissiveColor) * m.emissive_intensity;
    ret.Color = xyz(tx.grad(colorIdx)) * V3(m.color[0], m.color[1], m.color[2]);
    ret.Normal = ReconstructNormalFromXY(xyz(tx.grad(normalIdx)));
    ret.Transmission = m.transmission;
    ret.AttenuationColor = V3(m.attenuation_color[0], m.attenuation_color[1], m.attenuation_color[2]);
    ret.AttenuationDistance = m.attenuation_distance;
    ret.Eta = isHitFromInside ? m.ior : (1.0f / m.ior);
    vec3 specular = xyz(tx.grad(m.specular_idx)) * V3(m.specular[0], m.specular[1], m.specular[2]);
    float glossiness = tx.grad(m.glossiness_idx).w * m.glossiness;
    ret.Roughness = 1.0f - glossiness;
    const vec3 diff = max(specular - 0.04f, 0.0f) / ((ret.Color - 0.04f) + 0.00001f);
    ret.Metalness = (diff.x + diff.y + diff.z) / 3.0f;
    return ret;
}

/* PT/Shaders/material.glsl:144-171 (flags is always 0 in closestHit.rchit:102) */
MaterialSample sampleMaterial(const pto_scene &s, uint32_t materialId, vec2 texCoords, vec4 derivatives,
                              bool isHitFromInside, bool flipNormalY, Counters &c, uint32_t flags = 0)
{
    uint32_t materialType;
    const uint32_t materialIndex = unpackMaterialId(materialId, materialType);
    TexCtx tx { s, texCoords, V2(derivatives.x, derivatives.y), V2(derivatives.z, derivatives.w), {} };
    MaterialSample ret;
    switch (materialType)
    {
    case PT_MATERIAL_METALLIC_ROUGHNESS:
        ret = sampleMaterialMR(s.mr[materialIndex], tx, isHitFromInside, flags);
        break;
    case PT_MATERIAL_SPECULAR_GLOSSINESS:
        ret = sampleMaterialSG(s.sg[materialIndex], tx, isHitFromInside, flags);
        break;
    case PT_MATERIAL_PHONG:
        ret = sampleMaterialSG(s.phong[materialIndex], tx, isHitFromInside, flags);
        break;
    default:
        /* the GLSL leaves the other members undefined; zero them here */
        std::memset(&ret, 0, sizeof(ret));
        ret.Color = V3(1.0f, 0.0f, 0.0f);
        ret.EmissiveColor = V3(1.0f, 0.0f, 0.0f);
        break;
    }
    if (flipNormalY)
        ret.Normal.y *= -1.0f;
    c.texels += tx.tc.n;
    return ret;
}

/* ========================================================================= */
/* sampling.glsl                                                             */
/* ========================================================================= */

const float DirectionalLightDistance = 100000.0f; /* PT/Shaders/sampling.glsl:3 */

struct LightSample
{
    vec3 Direction;
    float Distance;
    vec3 Color;
    float Attenuation;
};

/* PT/Shaders/sampling.glsl:25-56 */
LightSample sampleLight(const pto_scene &s, vec3 u, vec3 position, float &pdf)
{
    const uint32_t u_LightCount = (uint32_t)s.pointLights.size();
    uint32_t lightIndex = (uint32_t)(u.x * (float)(u_LightCount + 1));
    pdf = 1.0f / (float)(u_LightCount + 1);
    LightSample ret;
    if (lightIndex >= u_LightCount)
    {
        const vec2 dp = sampleUniformDiskConcentric(V2(u.y, u.z));
        vec3 diskPoint = V3(dp.x, dp.y, 0.0f) * 0.001f;
        vec3 direction =
            normalize(V3(s.directional.direction[0], s.directional.direction[1], s.directional.direction[2]));
        ret.Direction = normalize(direction + computeTangentSpace(direction) * diskPoint);
        ret.Color = V3(s.directional.color[0], s.directional.color[1], s.directional.color[2]);
        ret.Distance = DirectionalLightDistance;
        ret.Attenuation = 1.0f;
        return ret;
    }
    const pt_point_light &light = s.pointLights[lightIndex];
    const vec3 lightPosition = V3(light.position[0], light.position[1], light.position[2]);
    const vec2 dp = sampleUniformDiskConcentric(V2(u.y, u.z));
    vec3 diskPoint = V3(dp.x, dp.y, 0.0f) * 0.1f;
    vec3 direction = normalize(position - lightPosition);
    vec3 newPosition = lightPosition + computeTangentSpace(direction) * diskPoint;
    ret.Distance = distance(position, newPosition);
    ret.Direction = normalize(position - newPosition);
    ret.Color = V3(light.color[0], light.color[1], light.color[2]);
    const float attenuation = 1.0f / (light.attenuation_constant + ret.Distance * light.attenuation_linear +
                                      ret.Distance * ret.Distance * light.attenuation_quadratic);
    ret.Attenuation = clamp(attenuation, 0.0f, 1.0f);
    return ret;
}

/* ========================================================================= */
/* Payload + stages                                                          */
/* ========================================================================= */

/* Shaders::Payload, PT/Shaders/ShaderRendererTypes.incl:101-118 */
struct Payload
{
    vec3 Position;
    vec3 Direction;
    float MaxRoughness;
    vec3 Bsdf;
    float Pdf;
    vec3 Emissive;
    uint32_t RngState;
    vec3 DirectLight;
    float DirectLightPdf; /* DecalDist   */
    vec3 LightDirection;  /* DecalAlbedo */
    float LightDistance;  /* DecalAlpha  */
    vec4 RayDifferentials0, RayDifferentials1, RayDifferentials2;
};

/* texture(samplerCube, dir) outside a fragment stage (miss.rmiss:31): level 0, linear filter.
 * Face selection and (s, t) follow the Vulkan specification ("Cube Map Face Selection": the major
 * axis is the largest |component|, z winning ties over y over x; table of sc / tc per face).
 * PARITY UNPINNED (sampler hardware): the bilinear footprint is clamped to the face here, the
 * hardware filters seamlessly across face edges — a sub-texel difference along the 12 seams. */
vec4 sampleCube(const Texture faces[6], vec3 r)
{
    const float ax = std::fabs(r.x), ay = std::fabs(r.y), az = std::fabs(r.z);
    int face;
    float sc, tc, ma;
    if (az >= ax && az >= ay)
    {
        face = r.z < 0.0f ? 5 : 4;
        sc = r.z < 0.0f ? -r.x : r.x;
        tc = -r.y;
        ma = az;
    }
    else if (ay >= ax)
    {
        face = r.y < 0.0f ? 3 : 2;
        sc = r.x;
        tc = r.y < 0.0f ? -r.z : r.z;
        ma = ay;
    }
    else
    {
        face = r.x < 0.0f ? 1 : 0;
        sc = r.x < 0.0f ? r.z : -r.z;
        tc = -r.y;
        ma = ax;
    }
    const Texture &t = faces[face];
    const TexLevel &l = t.levels[0];
    float u = 0.5f * (sc / ma + 1.0f), v = 0.5f * (tc / ma + 1.0f);
    u = std::isfinite(u) ? u : 0.5f;
    v = std::isfinite(v) ? v : 0.5f;
    const float x = u * (float)l.w - 0.5f, y = v * (float)l.h - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int W = (int)l.w, H = (int)l.h;
    const int x0 = std::min(std::max((int)fx0, 0), W - 1), x1 = std::min(std::max((int)fx0 + 1, 0), W - 1);
    const int y0 = std::min(std::max((int)fy0, 0), H - 1), y1 = std::min(std::max((int)fy0 + 1, 0), H - 1);
    const vec4 t00 = t.texel(0, x0, y0), t10 = t.texel(0, x1, y0);
    const vec4 t01 = t.texel(0, x0, y1), t11 = t.texel(0, x1, y1);
    const vec4 top = t00 * (1.0f - fx) + t10 * fx;
    const vec4 bot = t01 * (1.0f - fx) + t11 * fx;
    return top * (1.0f - fy) + bot * fy;
}

/* PT/Shaders/miss.rmiss:16-39 */
void missShader(const pto_scene &s, const pt_render_params &p, vec3 rayDir, Payload &payload)
{
    if ((p.miss_flags & PT_MISS_FLAGS_SKYBOX_2D) != 0 && s.hasSky2D)
    {
        const vec3 dir = rayDir;
        const float longitude = std::atan2(dir.z, dir.x);
        const float latitude = std::asin(-dir.y);
        const vec2 texCoords = V2(longitude / 2.0f / PI + 0.5f, latitude / PI + 0.5f);
        payload.Emissive = xyz(textureLod0(s.sky2D, texCoords));
        payload.Emissive = hdrToLdr(payload.Emissive);
    }
    else if ((p.miss_flags & PT_MISS_FLAGS_SKYBOX_CUBE) != 0 && s.hasSkyCube)
        payload.Emissive = xyz(sampleCube(s.skyCube, rayDir));
    else
        payload.Emissive = V3(0.08f, 0.09f, 0.1f);
    payload.Pdf = -1.0f;
}

/* PT/Shaders/closestHit.rchit:52-161 */
void closestHitShader(const pto_scene &s, const pt_render_params &p, const HitInfo &hit, vec3 rayOriginWorld,
                      vec3 rayDirWorld, Payload &payload, Counters &c)
{
    const FlatTri &ft = s.tris[hit.tri];
    const pt_instance &inst = s.instances[ft.instance];
    const pt_mesh_record &sbt = meshRecordOf(s, ft);
    const pt_geometry &geom = s.geometries[sbt.geometry_index];
    const mat3x4 objectToWorld = objectToWorld3x4(inst);
    const uint32_t primitiveID = ft.primitive;
    const float rayTmax = hit.t;

    const vec3 barycentricCoords = computeBarycentricCoords(V2(hit.b1, hit.b2));

    Vertex v0 = getVertex(s, geom, primitiveID * 3);
    Vertex v1 = getVertex(s, geom, primitiveID * 3 + 1);
    Vertex v2 = getVertex(s, geom, primitiveID * 3 + 2);
    const Vertex originalVertex = interpolate(v0, v1, v2, barycentricCoords);
    Vertex vertex = transformVertex(s, originalVertex, sbt.transform_index, objectToWorld);

    v0 = transformVertex(s, v0, sbt.transform_index, objectToWorld);
    v1 = transformVertex(s, v1, sbt.transform_index, objectToWorld);
    v2 = transformVertex(s, v2, sbt.transform_index, objectToWorld);

    vec3 edge1 = v1.Position - v0.Position;
    vec3 edge2 = v2.Position - v0.Position;
    vec3 geometricNormal = normalize(cross(edge1, edge2));

    const bool isHitFromInside = dot(geometricNormal, rayDirWorld) > 0.0f;
    if (isHitFromInside)
    {
        geometricNormal *= -1.0f;
        vertex.Normal *= -1.0f;
        vertex.Tangent *= -1.0f;
        vertex.Bitangent *= -1.0f;
    }

    const vec3 viewDir = rayDirWorld;
    (void)rayOriginWorld;

    vec3 dpdu, dpdv, dndu, dndv;
    computeDpnDuv(v0, v1, v2, vertex, dpdu, dpdv, dndu, dndv);

    vec3 rxOrigin = xyz(payload.RayDifferentials0);
    vec3 rxDirection = V3(payload.RayDifferentials0.w, payload.RayDifferentials1.x, payload.RayDifferentials1.y);
    vec3 ryOrigin = V3(payload.RayDifferentials1.z, payload.RayDifferentials1.w, payload.RayDifferentials2.x);
    vec3 ryDirection = V3(payload.RayDifferentials2.y, payload.RayDifferentials2.z, payload.RayDifferentials2.w);

    vec3 dpdx, dpdy;
    computeDpDxy(vertex.Position, rxOrigin, rxDirection, ryOrigin, ryDirection, vertex.Normal, dpdx, dpdy);

    const vec4 derivatives = computeDerivatives(dpdx, dpdy, dpdu, dpdv);

    const bool flipYNormal = (p.hit_flags & PT_HIT_FLAGS_DX_NORMAL_TEXTURES) != 0;
    MaterialSample material =
        sampleMaterial(s, sbt.material_id, vertex.TexCoords, derivatives, isHitFromInside, flipYNormal, c);

    /* decals (Q9) */
    if (payload.DirectLightPdf != -1.0f && rayTmax > payload.DirectLightPdf)
        material.Color = mix(material.Color, payload.LightDirection, payload.LightDistance);

    payload.MaxRoughness = max(material.Roughness, payload.MaxRoughness);
    material.Roughness = max(payload.MaxRoughness, 0.01f);

    const mat3 geometryTBN = M3(vertex.Tangent, vertex.Bitangent, vertex.Normal);
    const vec3 N = normalize(vertex.Normal + geometryTBN * material.Normal);
    const mat3 TBN = computeTangentSpace(N);
    const mat3 invTBN = inverse(TBN);
    const vec3 V = normalize(invTBN * normalize(-rayDirWorld));

    uint32_t rngState = payload.RngState;

    BSDFSample bsdf = sampleBSDF(material, V, rngState);

    if (isHitFromInside)
    {
        bsdf.Color.x *= std::pow(material.AttenuationColor.x, rayTmax / material.AttenuationDistance);
        bsdf.Color.y *= std::pow(material.AttenuationColor.y, rayTmax / material.AttenuationDistance);
        bsdf.Color.z *= std::pow(material.AttenuationColor.z, rayTmax / material.AttenuationDistance);
    }

    const bool isRefracted = bsdf.Direction.z < 0.0f;

    const vec3 rayOrigin = offsetRayOriginShadowTerminator(vertex, v0, v1, v2, barycentricCoords, isRefracted);

    float lightPdf, lightSmplPdf;
    const float l0 = rnd(rngState);
    const float l1 = rnd(rngState);
    const float l2 = rnd(rngState);
    LightSample light = sampleLight(s, V3(l0, l1, l2), rayOrigin, lightPdf);
    const vec3 L = normalize(invTBN * -light.Direction);
    const vec3 lightBsdf = evaluateBSDF(material, V, L, lightSmplPdf);

    payload.Direction = normalize(TBN * bsdf.Direction);
    if (isRefracted)
        payload.Position = offsetRayOriginSelfIntersection(vertex.Position, -geometricNormal);
    else
        payload.Position = rayOrigin;
    payload.Bsdf = bsdf.Color;
    payload.Pdf = bsdf.Pdf;
    payload.Emissive = material.EmissiveColor;
    payload.RngState = rngState;
    payload.DirectLight = light.Color * light.Attenuation * lightBsdf;
    payload.DirectLightPdf = lightPdf;
    payload.LightDirection = light.Direction;
    payload.LightDistance = light.Distance;

    if (isRefracted)
        computeRefractedDifferentialRays(derivatives, vertex.Normal, rayOrigin, -viewDir, payload.Direction, dndu, dndv,
                                         material.Eta, rxOrigin, rxDirection, ryOrigin, ryDirection);
    else
        computeReflectedDifferentialRays(derivatives, vertex.Normal, rayOrigin, -viewDir, payload.Direction, dndu, dndv,
                                         rxOrigin, rxDirection, ryOrigin, ryDirection);

    payload.RayDifferentials0 = V4(rxOrigin, rxDirection.x);
    payload.RayDifferentials1 = V4(rxDirection.y, rxDirection.z, ryOrigin.x, ryOrigin.y);
    payload.RayDifferentials2 = V4(ryOrigin.z, ryDirection.x, ryDirection.y, ryDirection.z);
}

/* traceRayEXT(... PrimaryRayHitGroupIndex ...) of raygen.rgen:68 */
void tracePrimaryType(const pto_scene &s, const pt_render_params &p, const Ray &ray, Payload &payload, Counters &c)
{
    const HitInfo hit = traceClosest(s, ray.Origin, ray.Direction, ray.tmin, ray.tmax, c);
    /* the any-hit stage wrote its decal record into the payload before closest-hit / miss run */
    if (hit.decalDist != -1.0f)
    {
        payload.LightDirection = hit.decalColor;
        payload.LightDistance = hit.decalAlpha;
        payload.DirectLightPdf = hit.decalDist;
    }
    if (hit.tri == PT_NO_HIT)
        missShader(s, p, ray.Direction, payload);
    else
        closestHitShader(s, p, hit, ray.Origin, ray.Direction, payload, c);
}

/* PT/Shaders/raygen.rgen:22-34 */
bool checkOccluded(const pto_scene &s, vec3 lightDir, vec3 position, float dist, Counters &c)
{
    vec3 direction = -normalize(lightDir);
    float tmin = 0.00001f;
    float tmax = dist;
    return traceOccluded(s, position, direction, tmin, tmax, c);
}

const uint32_t kMaxRestarts = 1024; /* raygen.rgen:99-112 can spin forever; the oracle and the core both cap it */

/* PT/Shaders/raygen.rgen:36-118, one invocation (SampleCount samples of one pixel) */
vec3 raygenPixel(const pto_scene &s, const pt_render_params &p, uint32_t px, uint32_t py, uint32_t width, uint32_t height,
                 uint32_t sampleCount, uint32_t totalSamples, Counters &c)
{
    const mat4 ViewInverse = toMat4(p.view_inverse);
    const mat4 ProjInverse = toMat4(p.proj_inverse);
    uint32_t rngState = initRng(px, py, width, totalSamples);
    vec3 radiance = V3(0.0f);
    Payload payload;
    std::memset(&payload, 0, sizeof(payload));
    uint32_t restarts = 0;

    for (int smpl = 0; smpl < (int)sampleCount; smpl++)
    {
        vec3 throughput = V3(1.0f);
        const float ux = rnd(rngState);
        const float uy = rnd(rngState);
        vec2 u = V2(ux, uy);
        Ray ray, rx, ry;
        const vec2 pixel = V2((float)px, (float)py), resolution = V2((float)width, (float)height);
        if (p.lens_radius > 0)
        {
            const float u2x = rnd(rngState);
            const float u2y = rnd(rngState);
            ray = constructPrimaryRayLens(pixel, resolution, ViewInverse, ProjInverse, u, V2(u2x, u2y), p.lens_radius,
                                          p.focal_distance, rx, ry);
        }
        else
            ray = constructPrimaryRay(pixel, resolution, ViewInverse, ProjInverse, u, rx, ry);

        payload.RayDifferentials0 = V4(rx.Origin, rx.Direction.x);
        payload.RayDifferentials1 = V4(rx.Direction.y, rx.Direction.z, ry.Origin.x, ry.Origin.y);
        payload.RayDifferentials2 = V4(ry.Origin.z, ry.Direction.x, ry.Direction.y, ry.Direction.z);
        payload.MaxRoughness = 0.0f;

        for (uint32_t bounce = 0; bounce < p.bounce_count; bounce++)
        {
            payload.RngState = rngState;
            payload.DirectLightPdf = -1.0f;
            payload.LightDirection = V3(0.0f);
            payload.LightDistance = 0.0f;
            tracePrimaryType(s, p, ray, payload, c);
            rngState = payload.RngState;

            if (payload.Pdf == -1.0f)
            {
                radiance += throughput * payload.Emissive;
                break;
            }
            radiance += throughput * payload.Emissive;

            if (payload.DirectLightPdf > 0.0f)
                if (!checkOccluded(s, payload.LightDirection, payload.Position, payload.LightDistance, c))
                    radiance += throughput * payload.DirectLight / payload.DirectLightPdf;

            if (payload.Pdf > 0.001f)
                throughput *= payload.Bsdf / payload.Pdf;

            const float prob = min(maxComponent(throughput), 1.0f);
            if (prob < 0.001f)
                break;
            if (prob < rnd(rngState))
                break;
            throughput /= prob;

            ray.Origin = payload.Position;
            ray.Direction = payload.Direction;
        }
        c.samples++;

        if (isnan(radiance.x) || isnan(radiance.y) || isnan(radiance.z) || isinf(radiance.x) || isinf(radiance.y) ||
            isinf(radiance.z))
        {
            radiance = V3(0.0f);
            if (++restarts > kMaxRestarts)
                break;
            c.restarts++;
            smpl = -1;
            continue;
        }
    }
    return radiance;
}

pt_hit toPtHit(const pto_scene &s, const HitInfo &h)
{
    pt_hit r;
    if (h.tri == PT_NO_HIT)
    {
        r.instance = r.geometry = r.primitive = PT_NO_HIT;
        r.t = r.u = r.v = 0.0f;
        return r;
    }
    const FlatTri &t = s.tris[h.tri];
    r.instance = t.instance;
    r.geometry = t.geometry;
    r.primitive = t.primitive;
    r.t = h.t;
    r.u = h.b1;
    r.v = h.b2;
    return r;
}

bool inTiles(uint32_t x, uint32_t y, const pt_tile *tiles, uint32_t n)
{
    if (!tiles)
        return true;
    for (uint32_t i = 0; i < n; i++)
        if (x >= tiles[i].x0 && x < tiles[i].x1 && y >= tiles[i].y0 && y < tiles[i].y1)
            return true;
    return false;
}

MaterialSample materialFromFloats(const float *f)
{
    MaterialSample m;
    m.EmissiveColor = V3(f[0], f[1], f[2]);
    m.Color = V3(f[3], f[4], f[5]);
    m.Normal = V3(f[6], f[7], f[8]);
    m.Roughness = f[9];
    m.Metalness = f[10];
    m.Transmission = f[11];
    m.Eta = f[12];
    m.AttenuationColor = V3(f[13], f[14], f[15]);
    m.AttenuationDistance = f[16];
    return m;
}

const uint32_t kTestIn[PT_TEST_MODE_COUNT] = PT_TEST_INPUT_STRIDES;
const uint32_t kTestOut[PT_TEST_MODE_COUNT] = PT_TEST_OUTPUT_STRIDES;

/* (position.xyz, uv.xy, normal.xyz) record of PT_TEST_DPN_DUV */
Vertex vertexPUN(const float *a)
{
    Vertex v;
    v.Position = V3(a[0], a[1], a[2]);
    v.TexCoords = V2(a[3], a[4]);
    v.Normal = V3(a[5], a[6], a[7]);
    v.Tangent = V3(0.0f);
    v.Bitangent = V3(0.0f);
    return v;
}

} // namespace

/* ========================================================================= */
/* Debug pipeline: Debug/debugRaygen.rgen, debugClosestHit.rchit,            */
/* debugAnyhit.rahit, debugMiss.rmiss                     */
/* ========================================================================= */

/* tracing.glsl:151-161 */
float computeLod(vec4 derivatives)
{
    const float dudx = derivatives.x, dvdx = derivatives.y, dudy = derivatives.z, dvdy = derivatives.w;
    const float sx = std::sqrt(dudx * dudx + dvdx * dvdx);
    const float sy = std::sqrt(dudy * dudy + dvdy * dvdy);
    const float smax = max(sx, sy);
    return smax == 0.0f ? 0.0f : std::log2(smax);
}

/* debugClosestHit.rchit:141-161 */
vec3 getRandomColor(uint32_t x)
{
    x *= 0x1eca7d79u;
    x ^= x >> 20;
    x = (x << 8) | (x >> 24);
    x = ~x;
    x ^= x << 5;
    x += 0x10afe4e7u;
    return V3((float)((x & 0xff000000u) >> 24) / 255.0f, (float)((x & 0x00ff0000u) >> 16) / 255.0f,
              (float)((x & 0x0000ff00u) >> 8) / 255.0f);
}

/* debugClosestHit.rchit:70-139: the raster-style Cook-Torrance term of the debug view */
vec3 debugLightContribution(vec3 lightDir, vec3 lightColor, float attenuation, vec3 V, vec3 N, vec3 color, float roughness,
                            float metalness)
{
    const vec3 L = -normalize(lightDir);
    const vec3 H = normalize(V + L);
    const vec3 radiance = lightColor * attenuation;
    vec3 F0 = V3(0.04f, 0.04f, 0.04f);
    F0 = mix(F0, color, metalness);
    /* DDDDistributionGGX */
    const float a = roughness * roughness, a2 = a * a;
    const float NdotH = max(dot(N, H), 0.0f), NdotH2 = NdotH * NdotH;
    float denomD = NdotH2 * (a2 - 1.0f) + 1.0f;
    denomD = PI * denomD * denomD;
    const float NDF = a2 / max(denomD, 0.0001f);
    /* DDDGeometrySmith */
    const float NdotV = max(dot(N, V), 0.0f), NdotL = max(dot(N, L), 0.0f);
    const float rr = roughness + 1.0f, k = (rr * rr) / 8.0f;
    const float ggx2 = NdotV / (NdotV * (1.0f - k) + k), ggx1 = NdotL / (NdotL * (1.0f - k) + k);
    const float G = ggx1 * ggx2;
    /* DDDfresnelSchlick */
    const float cosTheta = max(dot(H, V), 0.0f);
    const vec3 F = F0 + (V3(1.0f, 1.0f, 1.0f) - F0) * std::pow(clamp(1.0f - cosTheta, 0.0f, 1.0f), 5.0f);
    const vec3 numerator = F * (NDF * G);
    const float denominator = 4.0f * max(dot(N, V), 0.0f) * max(dot(N, L), 0.0f);
    const vec3 specular = numerator / max(denominator, 0.0001f);
    vec3 kD = V3(1.0f, 1.0f, 1.0f) - F;
    kD = kD * (1.0f - metalness);
    return (kD * color / PI + specular) * radiance * NdotL;
}

vec4 debugPixel(const pto_scene &s, const pt_render_params &p, const pt_debug_params &dbg, uint32_t x, uint32_t y,
                uint32_t width, uint32_t height, Counters &c)
{
    Ray rx, ry;
    const Ray ray = constructPrimaryRay(V2((float)x, (float)y), V2((float)width, (float)height), toMat4(p.view_inverse),
                                        toMat4(p.proj_inverse), V2(0.5f, 0.5f), rx, ry);
    const bool forceOpaque = (dbg.raygen_flags & PT_DEBUG_RAYGEN_FORCE_OPAQUE) != 0;
    const bool cull = (dbg.raygen_flags & PT_DEBUG_RAYGEN_CULL_BACK_FACES) != 0;
    const HitInfo hit = traceClosest(s, ray.Origin, ray.Direction, ray.tmin, ray.tmax, c, forceOpaque, nullptr, cull);
    if (hit.tri == PT_NO_HIT)
    {
        /* debugMiss.rmiss:17-36 (no hdrToLdr here) */
        if ((p.miss_flags & PT_MISS_FLAGS_SKYBOX_2D) != 0 && s.hasSky2D)
        {
            const vec3 dir = ray.Direction;
            const float longitude = std::atan2(dir.z, dir.x), latitude = std::asin(-dir.y);
            return V4(xyz(textureLod0(s.sky2D, V2(longitude / 2.0f / PI + 0.5f, latitude / PI + 0.5f))), 1.0f);
        }
        if ((p.miss_flags & PT_MISS_FLAGS_SKYBOX_CUBE) != 0 && s.hasSkyCube)
            return V4(xyz(sampleCube(s.skyCube, ray.Direction)), 1.0f);
        return V4(0.2f, 0.2f, 0.2f, 1.0f);
    }

    /* debugClosestHit.rchit:163-265 */
    const FlatTri &ft = s.tris[hit.tri];
    const pt_mesh_record &sbt = meshRecordOf(s, ft);
    const pt_geometry &geom = s.geometries[sbt.geometry_index];
    const mat3x4 objectToWorld = objectToWorld3x4(s.instances[ft.instance]);
    const vec3 barycentricCoords = computeBarycentricCoords(V2(hit.b1, hit.b2));
    Vertex v0 = getVertex(s, geom, ft.primitive * 3), v1 = getVertex(s, geom, ft.primitive * 3 + 1);
    Vertex v2 = getVertex(s, geom, ft.primitive * 3 + 2);
    const Vertex vertex = transformVertex(s, interpolate(v0, v1, v2, barycentricCoords), sbt.transform_index, objectToWorld);
    const vec3 origin = ray.Origin, viewDir = ray.Direction;
    v0 = transformVertex(s, v0, sbt.transform_index, objectToWorld);
    v1 = transformVertex(s, v1, sbt.transform_index, objectToWorld);
    v2 = transformVertex(s, v2, sbt.transform_index, objectToWorld);
    vec3 dpdu, dpdv, dndu, dndv;
    computeDpnDuv(v0, v1, v2, vertex, dpdu, dpdv, dndu, dndv);
    vec3 dpdx, dpdy;
    computeDpDxy(vertex.Position, origin, rx.Direction, origin, ry.Direction, vertex.Normal, dpdx, dpdy);
    const uint32_t flags = dbg.hit_group_flags;
    const vec4 derivatives = (flags & PT_DEBUG_HIT_DISABLE_MIP_MAPS) ? V4(0.0f, 0.0f, 0.0f, 0.0f) : computeDerivatives(dpdx, dpdy, dpdu, dpdv);
    const bool flipYNormal = (flags & PT_DEBUG_HIT_DX_NORMAL_TEXTURES) != 0;
    MaterialSample material = sampleMaterial(s, sbt.material_id, vertex.TexCoords, derivatives, false, flipYNormal, c, flags);
    if (hit.decalDist != -1.0f && hit.t > hit.decalDist)
        material.Color = mix(material.Color, hit.decalColor, hit.decalAlpha);
    const vec3 V = -normalize(viewDir);
    const mat3 TBN = M3(vertex.Tangent, vertex.Bitangent, vertex.Normal);
    const vec3 N = normalize(vertex.Normal + TBN * material.Normal);

    const float ambient = 0.1f;
    vec3 totalLight = material.Color * ambient + material.EmissiveColor;
    vec3 tmpu = vertex.Position - v0.Position, tmpv = vertex.Position - v1.Position, tmpw = vertex.Position - v2.Position;
    const float dotu = min(0.0f, dot(tmpu, v0.Normal)), dotv = min(0.0f, dot(tmpv, v1.Normal));
    const float dotw = min(0.0f, dot(tmpw, v2.Normal));
    tmpu = tmpu - v0.Normal * dotu;
    tmpv = tmpv - v1.Normal * dotv;
    tmpw = tmpw - v2.Normal * dotw;
    const vec3 Pp = vertex.Position + tmpu * barycentricCoords.x + tmpv * barycentricCoords.y + tmpw * barycentricCoords.z;
    const bool shadowsDisabled = (flags & PT_DEBUG_HIT_DISABLE_SHADOWS) != 0;
    auto checkOccluded = [&](vec3 lightDir, float dist) { return traceOccluded(s, Pp, -normalize(lightDir), 0.00001f, dist, c); };
    const vec3 dirDirection = V3(s.directional.direction[0], s.directional.direction[1], s.directional.direction[2]);
    const vec3 dirColor = V3(s.directional.color[0], s.directional.color[1], s.directional.color[2]);
    if (shadowsDisabled || !checkOccluded(dirDirection, DirectionalLightDistance))
        totalLight = totalLight + debugLightContribution(dirDirection, dirColor, 1.0f, V, N, material.Color, material.Roughness, material.Metalness);
    for (const pt_point_light &light : s.pointLights)
    {
        const vec3 lightDirection = Pp - V3(light.position[0], light.position[1], light.position[2]);
        const float dist = length(lightDirection);
        const float attenuation = 1.0f / (light.attenuation_constant + dist * light.attenuation_linear + dist * dist * light.attenuation_quadratic);
        if (shadowsDisabled || !checkOccluded(lightDirection, dist))
            totalLight = totalLight + debugLightContribution(lightDirection, V3(light.color[0], light.color[1], light.color[2]), attenuation, V, N,
                                                             material.Color, material.Roughness, material.Metalness);
    }
    switch (dbg.render_mode)
    {
    case PT_DEBUG_MODE_COLOR: return V4(totalLight, 1.0f);
    case PT_DEBUG_MODE_WORLD_POSITION: return V4(vertex.Position, 1.0f);
    case PT_DEBUG_MODE_NORMAL: return V4(N, 1.0f);
    case PT_DEBUG_MODE_TEXTURE_COORDS: return V4(vertex.TexCoords.x, vertex.TexCoords.y, 0.0f, 1.0f);
    case PT_DEBUG_MODE_MIPS: {
        const float v = 0.1f * computeLod(derivatives) + 1.0f;
        return V4(v, v, v, 1.0f);
    }
    case PT_DEBUG_MODE_GEOMETRY: return V4(getRandomColor(ft.geometry), 1.0f);
    case PT_DEBUG_MODE_PRIMITIVE: return V4(getRandomColor(ft.primitive), 1.0f);
    default: return V4(getRandomColor(ft.instance), 1.0f);
    }
}

/* ========================================================================= */
/* C ABI                                                                     */
/* ========================================================================= */

extern "C" {

pto_scene *pto_scene_create(const pt_scene_desc *d) { return pto_scene_create_limits(d, 4096, 0); }

pto_scene *pto_scene_create_limits(const pt_scene_desc *d, uint32_t max_texture_size, uint64_t texture_budget_bytes)
{
    if (!d)
        return nullptr;
    TextureLimits limits;
    limits.maxTextureSize = std::max(1u, max_texture_size);
    limits.budgetBytes = texture_budget_bytes;
    limits.textureCount = std::max(1u, d->texture_count);
    pto_scene *s = new pto_scene();
    s->vertices.assign(d->vertices, d->vertices + d->vertex_count);
    s->indices.assign(d->indices, d->indices + d->index_count);
    for (uint32_t i = 0; i < d->transform_count; i++)
    {
        const float *m = d->transforms + 12 * i;
        s->transforms.push_back(
            mat3x4 { { V4(m[0], m[1], m[2], m[3]), V4(m[4], m[5], m[6], m[7]), V4(m[8], m[9], m[10], m[11]) } });
    }
    s->geometries.assign(d->geometries, d->geometries + d->geometry_count);
    if (d->geometry_is_animated)
    {
        /* Renderer::RecordSkinningCommands (PT/Renderer/Renderer.cpp:854-890): the animated geometries'
         * vertices are skinned into an out-buffer the geometry table then points at (:353-372); here they
         * are appended to the static buffers and the animated geometries' offsets moved behind them */
        for (uint64_t i = 0; i < d->animated_vertex_count; i++)
            s->vertices.push_back(skinVertex(d->animated_vertices[i], d->bone_transforms, d->bone_count));
        s->indices.insert(s->indices.end(), d->animated_indices, d->animated_indices + d->animated_index_count);
        for (uint32_t g = 0; g < d->geometry_count; g++)
            if (d->geometry_is_animated[g])
            {
                s->geometries[g].vertex_offset += (uint32_t)d->vertex_count;
                s->geometries[g].index_offset += (uint32_t)d->index_count;
            }
    }
    s->meshRecords.assign(d->mesh_records, d->mesh_records + d->mesh_record_count);
    s->models.assign(d->models, d->models + d->model_count);
    s->instances.assign(d->instances, d->instances + d->instance_count);
    if (d->mr_material_count)
        s->mr.assign(d->mr_materials, d->mr_materials + d->mr_material_count);
    if (d->sg_material_count)
        s->sg.assign(d->sg_materials, d->sg_materials + d->sg_material_count);
    if (d->phong_material_count)
        s->phong.assign(d->phong_materials, d->phong_materials + d->phong_material_count);
    if (d->point_light_count)
        s->pointLights.assign(d->point_lights, d->point_lights + d->point_light_count);
    s->directional = d->directional_light;

    /* built-in textures, slots 0-8 (PT/Shaders/ShaderTypes.incl:18-26; texel values
     * ShaderRendererTypes.incl:49-56; sRGB rule TextureUploader.cpp:571-595 applied to the type) */
    s->textures.push_back(makeDefaultTexture(0xffffffffu, true));  /* 0 colour      */
    s->textures.push_back(makeDefaultTexture(0xffff8080u, false)); /* 1 normal      */
    s->textures.push_back(makeDefaultTexture(0xffffffffu, false)); /* 2 roughness   */
    s->textures.push_back(makeDefaultTexture(0xffffffffu, false)); /* 3 metalness   */
    s->textures.push_back(makeDefaultTexture(0x00000000u, true));  /* 4 emissive    */
    s->textures.push_back(makeDefaultTexture(0xffffffffu, true));  /* 5 specular    */
    s->textures.push_back(makeDefaultTexture(0x00000000u, false)); /* 6 glossiness  */
    s->textures.push_back(makeDefaultTexture(0x00000000u, false)); /* 7 shininess   */
    s->textures.push_back(makeDefaultTexture(0xffffffffu, true));  /* 8 placeholder */
    for (uint32_t i = 0; i < d->texture_count; i++)
        s->textures.push_back(makeTexture(d->textures[i], &limits));
    if (d->skybox_2d)
    {
        s->hasSky2D = true;
        s->sky2D = makeTexture(*d->skybox_2d);
    }
    if (d->skybox_cube)
    {
        s->hasSkyCube = true;
        for (int f = 0; f < 6; f++)
            s->skyCube[f] = makeTexture(d->skybox_cube[f]);
    }

    /* flatten TLAS -> BLAS -> geometry into world-space triangles
     * (PT/Renderer/AccelerationStructure.cpp:64-165, 260-301) */
    for (uint32_t ii = 0; ii < d->instance_count; ii++)
    {
        const pt_instance &inst = s->instances[ii];
        const pt_model &model = s->models[inst.model_index];
        const mat3x4 o2w = objectToWorld3x4(inst);
        for (uint32_t mi = 0; mi < model.mesh_count; mi++)
        {
            const pt_mesh_record &rec = s->meshRecords[model.mesh_offset + mi];
            const pt_geometry &g = s->geometries[rec.geometry_index];
            const mat3x4 transform = M4(s->transforms[rec.transform_index]) * o2w;
            for (uint32_t pi = 0; pi < g.index_length / 3; pi++)
            {
                FlatTri t;
                t.p0 = V4(getVertex(*s, g, pi * 3).Position, 1.0f) * transform;
                t.p1 = V4(getVertex(*s, g, pi * 3 + 1).Position, 1.0f) * transform;
                t.p2 = V4(getVertex(*s, g, pi * 3 + 2).Position, 1.0f) * transform;
                t.instance = ii;
                t.geometry = mi;
                t.primitive = pi;
                t.opaque = g.is_opaque;
                s->tris.push_back(t);
            }
        }
    }
    BvhBuilder(*s).build();
    return s;
}

void pto_scene_destroy(pto_scene *s) { delete s; }

int32_t pto_scene_set_sampler(pto_scene *s, uint32_t max_anisotropy)
{
    if (!s || max_anisotropy < 1 || max_anisotropy > 16)
        return PT_ERR_INVALID_ARGUMENT;
    s->maxAnisotropy = max_anisotropy;
    return PT_OK;
}

uint64_t pto_scene_triangle_count(const pto_scene *s) { return s ? s->tris.size() : 0; }

int32_t pto_render_frames(const pto_scene *s, const pt_render_params *p, uint32_t width, uint32_t height,
                          uint32_t first_sample, uint32_t frame_count, uint32_t samples_per_frame, const pt_tile *tiles,
                          uint32_t tile_count, float *accum, int32_t threads, pto_counters *out)
{
    if (!s || !p || !accum || samples_per_frame == 0)
        return PT_ERR_INVALID_ARGUMENT;
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1)
        nthreads = 1;
    std::vector<Counters> counters(nthreads);
    std::atomic<uint32_t> nextRow { 0 };
    auto worker = [&](int tid) {
        Counters &c = counters[tid];
        for (;;)
        {
            const uint32_t y = nextRow.fetch_add(1);
            if (y >= height)
                break;
            for (uint32_t x = 0; x < width; x++)
            {
                if (!inTiles(x, y, tiles, tile_count))
                    continue;
                float *px = accum + ((size_t)y * width + x) * 4;
                for (uint32_t f = 0; f < frame_count; f++)
                {
                    /* one frame = one vkCmdTraceRaysKHR: SampleCount = samples_per_frame, TotalSamples = samples
                     * accumulated so far (Renderer.cpp:1688-1700); raygen.rgen:115-117 */
                    const vec3 radiance = raygenPixel(*s, *p, x, y, width, height, samples_per_frame,
                                                      first_sample + f * samples_per_frame, c);
                    px[0] = radiance.x + px[0];
                    px[1] = radiance.y + px[1];
                    px[2] = radiance.z + px[2];
                    px[3] = 1.0f;
                }
            }
        }
    };
    if (nthreads == 1)
        worker(0);
    else
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < nthreads; i++)
            pool.emplace_back(worker, i);
        for (auto &t : pool)
            t.join();
    }
    if (out)
    {
        Counters total;
        for (auto &c : counters)
            total.add(c);
        out->rays_closest = total.rays_closest;
        out->rays_shadow = total.rays_shadow;
        out->samples = total.samples;
        out->hits = total.hits;
        out->box_tests_closest = total.box_c;
        out->tri_tests_closest = total.tri_c;
        out->box_tests_shadow = total.box_s;
        out->tri_tests_shadow = total.tri_s;
        out->alpha_tests_closest = total.alpha_c;
        out->alpha_tests_shadow = total.alpha_s;
        out->texel_fetches = total.texels;
        out->restarts = total.restarts;
    }
    return PT_OK;
}

int32_t pto_render(const pto_scene *s, const pt_render_params *p, uint32_t width, uint32_t height, uint32_t first_sample,
                   uint32_t sample_count, const pt_tile *tiles, uint32_t tile_count, float *accum, int32_t threads,
                   pto_counters *out)
{
    return pto_render_frames(s, p, width, height, first_sample, sample_count, 1, tiles, tile_count, accum, threads, out);
}

int32_t pto_trace_anyhit(const pto_scene *s, const float *org, const float *dir, float tmin, float tmax,
                         uint32_t terminate_on_first_hit, pto_anyhit_fn anyhit, void *ctx, pt_hit *out)
{
    if (!s || !org || !dir || !out)
        return 0;
    Counters c;
    AnyHitHook hook;
    hook.fn = anyhit;
    hook.ctx = ctx;
    const AnyHitHook *hp = anyhit ? &hook : nullptr;
    HitInfo h;
    if (terminate_on_first_hit)
        traceOccluded(*s, V3(org[0], org[1], org[2]), V3(dir[0], dir[1], dir[2]), tmin, tmax, c, hp, &h);
    else
        h = traceClosest(*s, V3(org[0], org[1], org[2]), V3(dir[0], dir[1], dir[2]), tmin, tmax, c, false, hp);
    *out = toPtHit(*s, h);
    return h.tri != PT_NO_HIT ? 1 : 0;
}

int32_t pto_trace_anyhit_flags(const pto_scene *s, const float *org, const float *dir, float tmin, float tmax,
                               uint32_t terminate_on_first_hit, uint32_t flags, pto_anyhit_fn anyhit, void *ctx, pt_hit *out)
{
    if (!s || !org || !dir || !out)
        return 0;
    if (terminate_on_first_hit)
        return pto_trace_anyhit(s, org, dir, tmin, tmax, terminate_on_first_hit, anyhit, ctx, out);
    Counters c;
    AnyHitHook hook;
    hook.fn = anyhit;
    hook.ctx = ctx;
    const HitInfo h = traceClosest(*s, V3(org[0], org[1], org[2]), V3(dir[0], dir[1], dir[2]), tmin, tmax, c, (flags & 1u) != 0,
                                   anyhit ? &hook : nullptr, (flags & 2u) != 0);
    *out = toPtHit(*s, h);
    return h.tri != PT_NO_HIT ? 1 : 0;
}

int32_t pto_sky_sample(const pto_scene *s, uint32_t kind, const float *in3, float *out4)
{
    if (!s || !in3 || !out4)
        return PT_ERR_INVALID_ARGUMENT;
    vec4 r = V4(0.0f, 0.0f, 0.0f, 0.0f);
    if (kind == 0 && s->hasSky2D)
        r = textureLod0(s->sky2D, V2(in3[0], in3[1]));
    else if (kind == 1 && s->hasSkyCube)
        r = sampleCube(s->skyCube, V3(in3[0], in3[1], in3[2]));
    else
        return PT_ERR_INVALID_ARGUMENT;
    out4[0] = r.x, out4[1] = r.y, out4[2] = r.z, out4[3] = r.w;
    return PT_OK;
}

int32_t pto_closest_hit(const pto_scene *s, const pt_render_params *p, uint32_t count, const pt_hit *hits, const float *rays6,
                        const float *payload_in, float *payload_out)
{
    if (!s || !p || !hits || !rays6 || !payload_in || !payload_out)
        return PT_ERR_INVALID_ARGUMENT;
    Counters c;
    for (uint32_t i = 0; i < count; i++)
    {
        /* find the flattened triangle of (instance, geometry, primitive) */
        HitInfo h;
        for (size_t ti = 0; ti < s->tris.size(); ti++)
        {
            const FlatTri &t = s->tris[ti];
            if (t.instance == hits[i].instance && t.geometry == hits[i].geometry && t.primitive == hits[i].primitive)
            {
                h.tri = (uint32_t)ti;
                break;
            }
        }
        if (h.tri == PT_NO_HIT)
            return PT_ERR_INVALID_ARGUMENT;
        h.t = hits[i].t;
        h.b1 = hits[i].u;
        h.b2 = hits[i].v;
        const float *f = payload_in + 36 * (size_t)i;
        /* Shaders::Payload, ShaderRendererTypes.incl:101-118, as 36 words */
        Payload pl;
        pl.Position = V3(f[0], f[1], f[2]);
        pl.Direction = V3(f[4], f[5], f[6]);
        pl.MaxRoughness = f[7];
        pl.Bsdf = V3(f[8], f[9], f[10]);
        pl.Pdf = f[11];
        pl.Emissive = V3(f[12], f[13], f[14]);
        pl.RngState = floatBitsToUint(f[15]);
        pl.DirectLight = V3(f[16], f[17], f[18]);
        pl.DirectLightPdf = f[19];
        pl.LightDirection = V3(f[20], f[21], f[22]);
        pl.LightDistance = f[23];
        pl.RayDifferentials0 = V4(f[24], f[25], f[26], f[27]);
        pl.RayDifferentials1 = V4(f[28], f[29], f[30], f[31]);
        pl.RayDifferentials2 = V4(f[32], f[33], f[34], f[35]);
        const float *r = rays6 + 6 * (size_t)i;
        closestHitShader(*s, *p, h, V3(r[0], r[1], r[2]), V3(r[3], r[4], r[5]), pl, c);
        float *o = payload_out + 36 * (size_t)i;
        o[0] = pl.Position.x, o[1] = pl.Position.y, o[2] = pl.Position.z, o[3] = f[3];
        o[4] = pl.Direction.x, o[5] = pl.Direction.y, o[6] = pl.Direction.z, o[7] = pl.MaxRoughness;
        o[8] = pl.Bsdf.x, o[9] = pl.Bsdf.y, o[10] = pl.Bsdf.z, o[11] = pl.Pdf;
        o[12] = pl.Emissive.x, o[13] = pl.Emissive.y, o[14] = pl.Emissive.z, o[15] = uintBitsToFloat(pl.RngState);
        o[16] = pl.DirectLight.x, o[17] = pl.DirectLight.y, o[18] = pl.DirectLight.z, o[19] = pl.DirectLightPdf;
        o[20] = pl.LightDirection.x, o[21] = pl.LightDirection.y, o[22] = pl.LightDirection.z, o[23] = pl.LightDistance;
        o[24] = pl.RayDifferentials0.x, o[25] = pl.RayDifferentials0.y, o[26] = pl.RayDifferentials0.z, o[27] = pl.RayDifferentials0.w;
        o[28] = pl.RayDifferentials1.x, o[29] = pl.RayDifferentials1.y, o[30] = pl.RayDifferentials1.z, o[31] = pl.RayDifferentials1.w;
        o[32] = pl.RayDifferentials2.x, o[33] = pl.RayDifferentials2.y, o[34] = pl.RayDifferentials2.z, o[35] = pl.RayDifferentials2.w;
    }
    return PT_OK;
}

int32_t pto_first_hit_aov(const pto_scene *s, const pt_render_params *p, uint32_t width, uint32_t height,
                          pt_hit *out_hits)
{
    if (!s || !p || !out_hits)
        return PT_ERR_INVALID_ARGUMENT;
    const mat4 ViewInverse = toMat4(p->view_inverse);
    const mat4 ProjInverse = toMat4(p->proj_inverse);
    /* rows are independent: all host threads */
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1 || (uint64_t)width * height < 65536)
        nthreads = 1;
    std::atomic<uint32_t> nextRow { 0 };
    auto worker = [&]() {
        Counters c;
        for (;;)
        {
            const uint32_t y = nextRow.fetch_add(1);
            if (y >= height)
                break;
            for (uint32_t x = 0; x < width; x++)
            {
                Ray rx, ry;
                /* PT/Shaders/ray.glsl:87-90: pixel centre */
                const Ray ray = constructPrimaryRay(V2((float)x, (float)y), V2((float)width, (float)height), ViewInverse,
                                                    ProjInverse, V2(0.5f, 0.5f), rx, ry);
                const HitInfo h = traceClosest(*s, ray.Origin, ray.Direction, ray.tmin, ray.tmax, c);
                out_hits[(size_t)y * width + x] = toPtHit(*s, h);
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
    return PT_OK;
}

int32_t pto_trace_closest(const pto_scene *s, const pt_ray *rays, uint64_t n, pt_hit *out_hits)
{
    if (!s || !rays || !out_hits)
        return PT_ERR_INVALID_ARGUMENT;
    Counters c;
    for (uint64_t i = 0; i < n; i++)
    {
        const pt_ray &r = rays[i];
        const HitInfo h = traceClosest(*s, V3(r.origin[0], r.origin[1], r.origin[2]),
                                       V3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, c);
        out_hits[i] = toPtHit(*s, h);
    }
    return PT_OK;
}

int32_t pto_trace_occlusion(const pto_scene *s, const pt_ray *rays, uint64_t n, uint8_t *out)
{
    if (!s || !rays || !out)
        return PT_ERR_INVALID_ARGUMENT;
    Counters c;
    for (uint64_t i = 0; i < n; i++)
    {
        const pt_ray &r = rays[i];
        out[i] = traceOccluded(*s, V3(r.origin[0], r.origin[1], r.origin[2]),
                               V3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, c)
                     ? 1
                     : 0;
    }
    return PT_OK;
}

int32_t pto_trace_closest_bruteforce_f64(const pto_scene *s, const pt_ray *rays, uint64_t n, pt_hit *out_hits,
                                         double *out_t)
{
    if (!s || !rays || !out_hits)
        return PT_ERR_INVALID_ARGUMENT;
    for (uint64_t i = 0; i < n; i++)
    {
        const pt_ray &r = rays[i];
        const double o[3] = { r.origin[0], r.origin[1], r.origin[2] };
        const double d[3] = { r.direction[0], r.direction[1], r.direction[2] };
        double best = r.tmax;
        HitInfo h;
        double bu = 0, bv = 0;
        for (size_t ti = 0; ti < s->tris.size(); ti++)
        {
            const FlatTri &t = s->tris[ti];
            const double p0[3] = { t.p0.x, t.p0.y, t.p0.z };
            const double e1[3] = { t.p1.x - p0[0], t.p1.y - p0[1], t.p1.z - p0[2] };
            const double e2[3] = { t.p2.x - p0[0], t.p2.y - p0[1], t.p2.z - p0[2] };
            const double pv[3] = { d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0] };
            const double det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
            if (det == 0.0)
                continue;
            const double inv = 1.0 / det;
            const double tv[3] = { o[0] - p0[0], o[1] - p0[1], o[2] - p0[2] };
            const double u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
            if (u < 0.0 || u > 1.0)
                continue;
            const double qv[3] = { tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2],
                                   tv[0] * e1[1] - tv[1] * e1[0] };
            const double v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
            if (v < 0.0 || u + v > 1.0)
                continue;
            const double tt = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
            if (!(tt > r.tmin && tt < best))
                continue;
            if (!t.opaque)
            {
                const vec4 color = anyHitColor(*s, t, (float)u, (float)v);
                if (color.w < 0.5f)
                    continue;
            }
            best = tt;
            h.tri = (uint32_t)ti;
            bu = u;
            bv = v;
        }
        h.t = (float)best;
        h.b1 = (float)bu;
        h.b2 = (float)bv;
        out_hits[i] = toPtHit(*s, h);
        if (out_t)
            out_t[i] = h.tri == PT_NO_HIT ? 0.0 : best;
    }
    return PT_OK;
}

int32_t pto_skin_vertices(const pt_animated_vertex *animated, const uint32_t *indices, uint64_t count,
                          const float *bone_transforms, uint32_t bone_count, pt_vertex *out)
{
    if (!animated || !indices || !bone_transforms || !out || bone_count == 0)
        return PT_ERR_INVALID_ARGUMENT;
    for (uint64_t i = 0; i < count; i++)
        out[i] = skinVertex(animated[indices[i]], bone_transforms, bone_count);
    return PT_OK;
}

int32_t pto_test_shading(uint32_t mode, const float *in, float *out, uint32_t count)
{
    if (mode >= PT_TEST_MODE_COUNT || !in || !out)
        return PT_ERR_INVALID_ARGUMENT;
    const uint32_t is = kTestIn[mode], os = kTestOut[mode];
    for (uint32_t i = 0; i < count; i++)
    {
        const float *a = in + (size_t)i * is;
        float *o = out + (size_t)i * os;
        switch (mode)
        {
        case PT_TEST_GGX_DISTRIBUTION:
            o[0] = GGXDistribution(V3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_LAMBDA:
            o[0] = Lambda(V3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_GGX_SMITH:
            o[0] = GGXSmith(V3(a[0], a[1], a[2]), a[3]);
            break;
        case PT_TEST_DIELECTRIC_FRESNEL:
            o[0] = DielectricFresnel(a[0], a[1]);
            break;
        case PT_TEST_SCHLICK_FRESNEL:
            o[0] = SchlickFresnel(a[0]);
            break;
        case PT_TEST_EVALUATE_REFLECTION: {
            float pdf;
            const vec3 f = EvaluateReflection(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[6], a[7], a[8]), a[9], pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_EVALUATE_REFRACTION: {
            float pdf;
            const vec3 f =
                EvaluateRefraction(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[6], a[7], a[8]), a[9], a[10], pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_SAMPLE_GGX: {
            const vec3 h = SampleGGX(V2(a[0], a[1]), V3(a[2], a[3], a[4]), a[5]);
            o[0] = h.x, o[1] = h.y, o[2] = h.z;
            break;
        }
        case PT_TEST_SAMPLE_LOBE_PDFS: {
            MaterialSample m;
            std::memset(&m, 0, sizeof(m));
            m.Metalness = a[0];
            m.Transmission = a[1];
            const LobePdfs p = sampleLobePdfs(m, a[2]);
            o[0] = p.Diffuse, o[1] = p.Glossy, o[2] = p.Metallic, o[3] = p.Transmissive;
            break;
        }
        case PT_TEST_EVALUATE_BSDF: {
            const MaterialSample m = materialFromFloats(a);
            float pdf;
            const vec3 f = evaluateBSDF(m, V3(a[17], a[18], a[19]), V3(a[20], a[21], a[22]), pdf);
            o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
            break;
        }
        case PT_TEST_SAMPLE_BSDF: {
            const MaterialSample m = materialFromFloats(a);
            uint32_t rng = floatBitsToUint(a[20]);
            const BSDFSample b = sampleBSDF(m, V3(a[17], a[18], a[19]), rng);
            o[0] = b.Direction.x, o[1] = b.Direction.y, o[2] = b.Direction.z, o[3] = b.Pdf;
            o[4] = b.Color.x, o[5] = b.Color.y, o[6] = b.Color.z, o[7] = uintBitsToFloat(rng);
            break;
        }
        case PT_TEST_RNG: {
            uint32_t st = initRng(floatBitsToUint(a[0]), floatBitsToUint(a[1]), floatBitsToUint(a[2]), floatBitsToUint(a[3]));
            o[0] = uintBitsToFloat(st);
            for (int k = 0; k < 4; k++)
            {
                const float f = rnd(st);
                o[1 + k] = uintBitsToFloat(st);
                o[5 + k] = f;
            }
            break;
        }
        case PT_TEST_PRIMARY_RAY: {
            const vec2 pixel = V2((float)floatBitsToUint(a[0]), (float)floatBitsToUint(a[1]));
            const vec2 res = V2((float)floatBitsToUint(a[2]), (float)floatBitsToUint(a[3]));
            const mat4 vi = toMat4(a + 10), pi = toMat4(a + 26);
            Ray rx, ry, r;
            if (a[8] > 0)
                r = constructPrimaryRayLens(pixel, res, vi, pi, V2(a[4], a[5]), V2(a[6], a[7]), a[8], a[9], rx, ry);
            else
                r = constructPrimaryRay(pixel, res, vi, pi, V2(a[4], a[5]), rx, ry);
            const Ray *rs[3] = { &r, &rx, &ry };
            for (int k = 0; k < 3; k++)
            {
                o[k * 6 + 0] = rs[k]->Origin.x, o[k * 6 + 1] = rs[k]->Origin.y, o[k * 6 + 2] = rs[k]->Origin.z;
                o[k * 6 + 3] = rs[k]->Direction.x, o[k * 6 + 4] = rs[k]->Direction.y, o[k * 6 + 5] = rs[k]->Direction.z;
            }
            break;
        }
        case PT_TEST_OFFSET_SELF_INTERSECTION: {
            const vec3 r = offsetRayOriginSelfIntersection(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_CONCENTRIC_DISK: {
            const vec2 r = sampleUniformDiskConcentric(V2(a[0], a[1]));
            o[0] = r.x, o[1] = r.y;
            break;
        }
        case PT_TEST_TANGENT_SPACE: {
            const mat3 m = computeTangentSpace(V3(a[0], a[1], a[2]));
            for (int k = 0; k < 3; k++)
                o[k * 3] = m.c[k].x, o[k * 3 + 1] = m.c[k].y, o[k * 3 + 2] = m.c[k].z;
            break;
        }
        case PT_TEST_DPN_DUV: {
            const Vertex v0 = vertexPUN(a), v1 = vertexPUN(a + 8), v2 = vertexPUN(a + 16);
            Vertex vertex = vertexPUN(a);
            vertex.Tangent = V3(a[24], a[25], a[26]);
            vertex.Bitangent = V3(a[27], a[28], a[29]);
            vec3 r[4];
            computeDpnDuv(v0, v1, v2, vertex, r[0], r[1], r[2], r[3]);
            for (int k = 0; k < 4; k++)
                o[k * 3] = r[k].x, o[k * 3 + 1] = r[k].y, o[k * 3 + 2] = r[k].z;
            break;
        }
        case PT_TEST_DP_DXY: {
            vec3 dpdx, dpdy;
            computeDpDxy(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[6], a[7], a[8]), V3(a[9], a[10], a[11]),
                         V3(a[12], a[13], a[14]), V3(a[15], a[16], a[17]), dpdx, dpdy);
            o[0] = dpdx.x, o[1] = dpdx.y, o[2] = dpdx.z, o[3] = dpdy.x, o[4] = dpdy.y, o[5] = dpdy.z;
            break;
        }
        case PT_TEST_DERIVATIVES: {
            const vec4 r = computeDerivatives(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[6], a[7], a[8]),
                                              V3(a[9], a[10], a[11]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z, o[3] = r.w;
            break;
        }
        case PT_TEST_REFLECTED_DIFFERENTIALS:
        case PT_TEST_REFRACTED_DIFFERENTIALS: {
            const bool refr = mode == PT_TEST_REFRACTED_DIFFERENTIALS;
            const float *b = a + 22 + (refr ? 1 : 0);
            vec3 r[4] = { V3(b[0], b[1], b[2]), V3(b[3], b[4], b[5]), V3(b[6], b[7], b[8]), V3(b[9], b[10], b[11]) };
            const vec4 der = V4(a[0], a[1], a[2], a[3]);
            const vec3 n = V3(a[4], a[5], a[6]), pp = V3(a[7], a[8], a[9]), viewDir = V3(a[10], a[11], a[12]);
            const vec3 outDir = V3(a[13], a[14], a[15]), dndu = V3(a[16], a[17], a[18]), dndv = V3(a[19], a[20], a[21]);
            if (refr)
                computeRefractedDifferentialRays(der, n, pp, viewDir, outDir, dndu, dndv, a[22], r[0], r[1], r[2], r[3]);
            else
                computeReflectedDifferentialRays(der, n, pp, viewDir, outDir, dndu, dndv, r[0], r[1], r[2], r[3]);
            for (int k = 0; k < 4; k++)
                o[k * 3] = r[k].x, o[k * 3 + 1] = r[k].y, o[k * 3 + 2] = r[k].z;
            break;
        }
        case PT_TEST_SHADOW_TERMINATOR: {
            Vertex vertex = vertexPUN(a), v0 = vertex, v1 = vertex, v2 = vertex;
            vertex.Position = V3(a[0], a[1], a[2]);
            v0.Position = V3(a[3], a[4], a[5]), v0.Normal = V3(a[6], a[7], a[8]);
            v1.Position = V3(a[9], a[10], a[11]), v1.Normal = V3(a[12], a[13], a[14]);
            v2.Position = V3(a[15], a[16], a[17]), v2.Normal = V3(a[18], a[19], a[20]);
            const vec3 r = offsetRayOriginShadowTerminator(vertex, v0, v1, v2, V3(a[21], a[22], a[23]), a[24] != 0.0f);
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_SAMPLE_LIGHT: {
            pto_scene sc;
            std::memset(&sc.directional, 0, sizeof(sc.directional));
            for (int k = 0; k < 3; k++)
                sc.directional.color[k] = a[6 + k], sc.directional.direction[k] = a[9 + k];
            if (floatBitsToUint(a[21]))
            {
                pt_point_light pl;
                std::memset(&pl, 0, sizeof(pl));
                for (int k = 0; k < 3; k++)
                    pl.color[k] = a[12 + k], pl.position[k] = a[15 + k];
                pl.attenuation_constant = a[18], pl.attenuation_linear = a[19], pl.attenuation_quadratic = a[20];
                sc.pointLights.push_back(pl);
            }
            float pdf;
            const LightSample l = sampleLight(sc, V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), pdf);
            o[0] = l.Direction.x, o[1] = l.Direction.y, o[2] = l.Direction.z, o[3] = l.Distance;
            o[4] = l.Color.x, o[5] = l.Color.y, o[6] = l.Color.z, o[7] = l.Attenuation, o[8] = pdf;
            break;
        }
        case PT_TEST_TRANSFORM_VERTEX: {
            pto_scene sc;
            Vertex v = vertexPUN(a);
            v.Position = V3(a[0], a[1], a[2]), v.Normal = V3(a[3], a[4], a[5]);
            v.Tangent = V3(a[6], a[7], a[8]), v.Bitangent = V3(a[9], a[10], a[11]);
            const float *m = a + 12, *w = a + 24;
            sc.transforms.push_back(mat3x4 { { V4(m[0], m[1], m[2], m[3]), V4(m[4], m[5], m[6], m[7]), V4(m[8], m[9], m[10], m[11]) } });
            const mat3x4 o2w = { { V4(w[0], w[1], w[2], w[3]), V4(w[4], w[5], w[6], w[7]), V4(w[8], w[9], w[10], w[11]) } };
            const Vertex r = transformVertex(sc, v, 0, o2w);
            o[0] = r.Position.x, o[1] = r.Position.y, o[2] = r.Position.z, o[3] = r.Normal.x, o[4] = r.Normal.y, o[5] = r.Normal.z;
            o[6] = r.Tangent.x, o[7] = r.Tangent.y, o[8] = r.Tangent.z, o[9] = r.Bitangent.x, o[10] = r.Bitangent.y,
            o[11] = r.Bitangent.z;
            break;
        }
        case PT_TEST_RECONSTRUCT_NORMAL: {
            const vec3 r = ReconstructNormalFromXY(V3(a[0], a[1], a[2]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        case PT_TEST_HDR_TO_LDR: {
            const vec3 r = hdrToLdr(V3(a[0], a[1], a[2]));
            o[0] = r.x, o[1] = r.y, o[2] = r.z;
            break;
        }
        }
    }
    return PT_OK;
}

int32_t pto_texture_info(const pto_scene *s, uint32_t slot, uint32_t *w, uint32_t *h, uint32_t *levels)
{
    if (!s || slot >= s->textures.size())
        return PT_ERR_INVALID_ARGUMENT;
    const Texture &t = s->textures[slot];
    if (w)
        *w = t.levels[0].w;
    if (h)
        *h = t.levels[0].h;
    if (levels)
        *levels = (uint32_t)t.levels.size();
    return PT_OK;
}

int32_t pto_texture_level(const pto_scene *s, uint32_t slot, uint32_t level, uint8_t *out)
{
    if (!s || slot >= s->textures.size() || !out)
        return PT_ERR_INVALID_ARGUMENT;
    const Texture &t = s->textures[slot];
    if (t.isFloat || level >= t.levels.size())
        return PT_ERR_INVALID_ARGUMENT;
    std::memcpy(out, t.levels[level].rgba8.data(), t.levels[level].rgba8.size());
    return PT_OK;
}

int32_t pto_texture_sample(const pto_scene *s, uint32_t slot, const float *in6, float *out4, uint32_t count,
                           int32_t use_grad)
{
    if (!s || slot >= s->textures.size() || !in6 || !out4)
        return PT_ERR_INVALID_ARGUMENT;
    const Texture &t = s->textures[slot];
    for (uint32_t i = 0; i < count; i++)
    {
        const float *a = in6 + (size_t)i * 6;
        const vec4 r = use_grad ? textureGrad(t, V2(a[0], a[1]), V2(a[2], a[3]), V2(a[4], a[5]), nullptr, s->maxAnisotropy)
                                : textureLod0(t, V2(a[0], a[1]));
        out4[i * 4] = r.x, out4[i * 4 + 1] = r.y, out4[i * 4 + 2] = r.z, out4[i * 4 + 3] = r.w;
    }
    return PT_OK;
}

int32_t pto_debug_render(const pto_scene *s, const pt_render_params *p, const pt_debug_params *dbg, uint32_t width,
                         uint32_t height, float *out_rgba)
{
    if (!s || !p || !dbg || !out_rgba || dbg->render_mode > PT_DEBUG_MODE_INSTANCE)
        return PT_ERR_INVALID_ARGUMENT;
    Counters c;
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++)
        {
            const vec4 v = debugPixel(*s, *p, *dbg, x, y, width, height, c);
            float *o = out_rgba + 4 * ((size_t)y * width + x);
            o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
        }
    return PT_OK;
}

} /* extern "C" */
