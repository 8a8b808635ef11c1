// unit_kernels.cu — pt_test_shading: runs the production __device__ functions one record per
// thread, the CUDA counterpart of the reference's shader unit-test harness
// (PTT/Shaders/testShading.comp, testBsdf.comp driven by PTT/TestRenderer.cpp:79-106).
#include "core_internal.h"
#include "shading.cuh"

#include <cstring>
#include <vector>

namespace pt
{

namespace
{

__constant__ uint32_t c_in[PT_TEST_MODE_COUNT] = PT_TEST_INPUT_STRIDES;
__constant__ uint32_t c_out[PT_TEST_MODE_COUNT] = PT_TEST_OUTPUT_STRIDES;
const uint32_t h_in[PT_TEST_MODE_COUNT] = PT_TEST_INPUT_STRIDES;
const uint32_t h_out[PT_TEST_MODE_COUNT] = PT_TEST_OUTPUT_STRIDES;

__device__ MaterialSample materialFromFloats(const float *f)
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

__global__ void k_test(uint32_t mode, const float *__restrict__ in, float *__restrict__ out, uint32_t count, LightBlock *scratch)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    const float *a = in + (size_t)i * c_in[mode];
    float *o = out + (size_t)i * c_out[mode];
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
        const LobePdfs p = sampleLobePdfs(a[0], a[1], a[2]);
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
        uint32_t rng = __float_as_uint(a[20]);
        const BSDFSample b = sampleBSDF(m, V3(a[17], a[18], a[19]), rng);
        o[0] = b.Direction.x, o[1] = b.Direction.y, o[2] = b.Direction.z, o[3] = b.Pdf;
        o[4] = b.Color.x, o[5] = b.Color.y, o[6] = b.Color.z, o[7] = __uint_as_float(rng);
        break;
    }
    case PT_TEST_RNG: {
        uint32_t st = initRng(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
        o[0] = __uint_as_float(st);
        for (int k = 0; k < 4; k++)
        {
            const float f = rnd(st);
            o[1 + k] = __uint_as_float(st);
            o[5 + k] = f;
        }
        break;
    }
    case PT_TEST_PRIMARY_RAY: {
        CameraMatrices cam;
        for (int k = 0; k < 16; k++)
        {
            cam.view[k] = a[10 + k];
            cam.proj[k] = a[26 + k];
        }
        const PrimaryRays r = constructPrimaryRay((float)__float_as_uint(a[0]), (float)__float_as_uint(a[1]),
                                                  (float)__float_as_uint(a[2]), (float)__float_as_uint(a[3]), cam,
                                                  V2(a[4], a[5]), V2(a[6], a[7]), a[8], a[9]);
        const vec3 dirs[3] = { r.direction, r.rxDirection, r.ryDirection };
        for (int k = 0; k < 3; k++)
        {
            o[k * 6 + 0] = r.origin.x, o[k * 6 + 1] = r.origin.y, o[k * 6 + 2] = r.origin.z;
            o[k * 6 + 3] = dirs[k].x, o[k * 6 + 4] = dirs[k].y, o[k * 6 + 5] = dirs[k].z;
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
        o[0] = m.c0.x, o[1] = m.c0.y, o[2] = m.c0.z;
        o[3] = m.c1.x, o[4] = m.c1.y, o[5] = m.c1.z;
        o[6] = m.c2.x, o[7] = m.c2.y, o[8] = m.c2.z;
        break;
    }
    case PT_TEST_DPN_DUV: {
        // records are (position, uv, normal) per corner; the production function takes the differences
        const vec3 p0 = V3(a[0], a[1], a[2]), p1 = V3(a[8], a[9], a[10]), p2 = V3(a[16], a[17], a[18]);
        const vec2 uv0 = V2(a[3], a[4]), uv1 = V2(a[11], a[12]), uv2 = V2(a[19], a[20]);
        const vec3 n0 = V3(a[5], a[6], a[7]), n1 = V3(a[13], a[14], a[15]), n2 = V3(a[21], a[22], a[23]);
        vec3 r[4];
        computeDpnDuv(p1 - p0, p2 - p0, n1 - n0, n2 - n0, uv1 - uv0, uv2 - uv0, V3(a[24], a[25], a[26]), V3(a[27], a[28], a[29]),
                      r[0], r[1], r[2], r[3]);
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
        const float4 r = computeDerivatives(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[6], a[7], a[8]), V3(a[9], a[10], a[11]));
        o[0] = r.x, o[1] = r.y, o[2] = r.z, o[3] = r.w;
        break;
    }
    case PT_TEST_REFLECTED_DIFFERENTIALS:
    case PT_TEST_REFRACTED_DIFFERENTIALS: {
        const bool refr = mode == PT_TEST_REFRACTED_DIFFERENTIALS;
        const float *b = a + 22 + (refr ? 1 : 0);
        RayDifferentials rd;
        rd.rxOrigin = V3(b[0], b[1], b[2]), rd.rxDirection = V3(b[3], b[4], b[5]);
        rd.ryOrigin = V3(b[6], b[7], b[8]), rd.ryDirection = V3(b[9], b[10], b[11]);
        propagateDifferentials(make_float4(a[0], a[1], a[2], a[3]), V3(a[4], a[5], a[6]), V3(a[7], a[8], a[9]),
                               V3(a[10], a[11], a[12]), V3(a[13], a[14], a[15]), V3(a[16], a[17], a[18]), V3(a[19], a[20], a[21]),
                               refr ? a[22] : 1.0f, refr, rd);
        const vec3 r[4] = { rd.rxOrigin, rd.rxDirection, rd.ryOrigin, rd.ryDirection };
        for (int k = 0; k < 4; k++)
            o[k * 3] = r[k].x, o[k * 3 + 1] = r[k].y, o[k * 3 + 2] = r[k].z;
        break;
    }
    case PT_TEST_SHADOW_TERMINATOR: {
        const vec3 r = offsetRayOriginShadowTerminator(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), V3(a[9], a[10], a[11]),
                                                       V3(a[15], a[16], a[17]), V3(a[6], a[7], a[8]), V3(a[12], a[13], a[14]),
                                                       V3(a[18], a[19], a[20]), V3(a[21], a[22], a[23]), a[24] != 0.0f);
        o[0] = r.x, o[1] = r.y, o[2] = r.z;
        break;
    }
    case PT_TEST_SAMPLE_LIGHT: {
        // the production function reads the light block through __ldg: testShading built one per record
        const LightBlock *lb = scratch + i;
        float pdf;
        const LightSample l = sampleLight(lb, V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5]), pdf);
        o[0] = l.Direction.x, o[1] = l.Direction.y, o[2] = l.Direction.z, o[3] = l.Distance;
        o[4] = l.Color.x, o[5] = l.Color.y, o[6] = l.Color.z, o[7] = l.Attenuation, o[8] = pdf;
        break;
    }
    case PT_TEST_TRANSFORM_VERTEX: {
        // a[12..23]: the (instance x mesh) matrix P and a[24..32]: its normal matrix N, both composed on the host by
        // the production flattenInstances arithmetic (testShading rewrites the record); then k_bake's arithmetic
        const float *P = a + 12, *N = a + 24;
        vec3 pos, nrm, tan, bit;
        float *pp = &pos.x, *pn = &nrm.x, *pt = &tan.x, *pb = &bit.x;
        for (int j = 0; j < 3; j++)
        {
            const float *Pj = P + j * 4;
            pp[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[0], Pj[0]), __fmul_rn(a[1], Pj[1])), __fmul_rn(a[2], Pj[2])), Pj[3]);
            pn[j] = N[j * 3 + 0] * a[3] + N[j * 3 + 1] * a[4] + N[j * 3 + 2] * a[5];
            pt[j] = Pj[0] * a[6] + Pj[1] * a[7] + Pj[2] * a[8];
            pb[j] = Pj[0] * a[9] + Pj[1] * a[10] + Pj[2] * a[11];
        }
        nrm = normalize(nrm), tan = normalize(tan), bit = normalize(bit);
        o[0] = pos.x, o[1] = pos.y, o[2] = pos.z, o[3] = nrm.x, o[4] = nrm.y, o[5] = nrm.z;
        o[6] = tan.x, o[7] = tan.y, o[8] = tan.z, o[9] = bit.x, o[10] = bit.y, o[11] = bit.z;
        break;
    }
    case PT_TEST_RECONSTRUCT_NORMAL: {
        const vec3 r = reconstructNormalFromXY(make_float4(a[0], a[1], a[2], 0.0f));
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

// texture(sampler, uv) (level 0: anyhit.rahit:51, miss.rmiss:27) / textureGrad(sampler, uv, ddx, ddy)
// (material.glsl:62-171) of one bindless slot, one record per thread
__global__ void k_test_texture(DeviceScene s, uint32_t slot, const float *__restrict__ in, float *__restrict__ out, uint32_t count,
                               int useGrad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    const float *a = in + (size_t)i * 6;
    const float4 r = useGrad ? textureGrad(s, slot, a[0], a[1], make_float4(a[2], a[3], a[4], a[5]))
                             : textureLod0(s, s.textures[slot], a[0], a[1]);
    out[(size_t)i * 4 + 0] = r.x, out[(size_t)i * 4 + 1] = r.y, out[(size_t)i * 4 + 2] = r.z, out[(size_t)i * 4 + 3] = r.w;
}

} // namespace

pt_status testTexture(Context *ctx, uint32_t slot, const float *in6, float *out4, uint32_t count, int32_t useGrad)
{
    if (!ctx->hasScene)
        return fail(ctx, PT_ERR_NO_SCENE, "pt_test_texture", "no scene uploaded");
    if (slot >= ctx->scene.textureCount || (count && (!in6 || !out4)))
        return fail(ctx, PT_ERR_INVALID_ARGUMENT, "pt_test_texture", "slot out of range or NULL buffer");
    if (count == 0)
        return PT_OK;
    float *dIn = nullptr, *dOut = nullptr;
    PT_CUDA_CHECK(ctx, cudaMalloc((void **)&dIn, (size_t)count * 24));
    cudaError_t err = cudaMalloc((void **)&dOut, (size_t)count * 16);
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(dIn, in6, (size_t)count * 24, cudaMemcpyHostToDevice, ctx->stream);
    if (err == cudaSuccess)
    {
        k_test_texture<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, slot, dIn, dOut, count, useGrad);
        err = cudaMemcpyAsync(out4, dOut, (size_t)count * 16, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (err == cudaSuccess)
        err = cudaStreamSynchronize(ctx->stream);
    cudaFree(dIn);
    cudaFree(dOut);
    PT_CUDA_CHECK(ctx, err);
    return PT_OK;
}

uint32_t testInputStride(uint32_t mode) { return mode < PT_TEST_MODE_COUNT ? h_in[mode] : 0; }
uint32_t testOutputStride(uint32_t mode) { return mode < PT_TEST_MODE_COUNT ? h_out[mode] : 0; }

pt_status testShading(Context *ctx, uint32_t mode, const float *input, float *output, uint32_t count)
{
    if (mode >= PT_TEST_MODE_COUNT || (count && (!input || !output)))
        return fail(ctx, PT_ERR_INVALID_ARGUMENT, "pt_test_shading", "bad mode or NULL buffer");
    if (count == 0)
        return PT_OK;
    const size_t inBytes = (size_t)count * h_in[mode] * 4, outBytes = (size_t)count * h_out[mode] * 4;
    std::vector<float> composed;
    if (mode == PT_TEST_TRANSFORM_VERTEX)
    {
        // the matrices are composed and inverted on the host in production (flattenInstances)
        composed.assign(input, input + (size_t)count * h_in[mode]);
        for (uint32_t i = 0; i < count; i++)
        {
            float *a = composed.data() + (size_t)i * h_in[mode];
            float P[12], N[9];
            testComposeTransform(a + 12, a + 24, P, N);
            std::memcpy(a + 12, P, sizeof(P));
            std::memcpy(a + 24, N, sizeof(N));
        }
        input = composed.data();
    }
    float *dIn = nullptr, *dOut = nullptr;
    LightBlock *dScratch = nullptr;
    PT_CUDA_CHECK(ctx, cudaMalloc((void **)&dIn, inBytes));
    cudaError_t err = cudaMalloc((void **)&dOut, outBytes);
    if (err == cudaSuccess && mode == PT_TEST_SAMPLE_LIGHT)
    {
        std::vector<LightBlock> blocks(count);
        for (uint32_t i = 0; i < count; i++)
        {
            const float *a = input + (size_t)i * h_in[mode];
            LightBlock &lb = blocks[i];
            std::memset(&lb, 0, sizeof(lb));
            uint32_t bits;
            std::memcpy(&bits, a + 21, 4);
            lb.count = bits ? 1u : 0u;
            lb.dirColor = make_float4(a[6], a[7], a[8], 0.0f);
            lb.dirDirection = make_float4(a[9], a[10], a[11], 0.0f);
            lb.point[0] = make_float4(a[12], a[13], a[14], 0.0f);
            lb.point[1] = make_float4(a[15], a[16], a[17], 0.0f);
            lb.point[2] = make_float4(a[18], a[19], a[20], 0.0f);
        }
        err = cudaMalloc((void **)&dScratch, (size_t)count * sizeof(LightBlock));
        if (err == cudaSuccess)
            err = cudaMemcpy(dScratch, blocks.data(), (size_t)count * sizeof(LightBlock), cudaMemcpyHostToDevice);
    }
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(dIn, input, inBytes, cudaMemcpyHostToDevice, ctx->stream);
    if (err == cudaSuccess)
    {
        k_test<<<(count + 63) / 64, 64, 0, ctx->stream>>>(mode, dIn, dOut, count, dScratch);
        err = cudaMemcpyAsync(output, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (err == cudaSuccess)
        err = cudaStreamSynchronize(ctx->stream);
    cudaFree(dIn);
    cudaFree(dOut);
    cudaFree(dScratch);
    PT_CUDA_CHECK(ctx, err);
    return PT_OK;
}

} // namespace pt
