// core_internal.h — host-side state of a pt_context (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/pt_core.h"
#include "scene.cuh"

#define PT_CUDA_CHECK(ctx, expr)                                                                                      \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t err__ = (expr);                                                                                   \
        if (err__ != cudaSuccess)                                                                                     \
            return pt::fail(ctx, err__ == cudaErrorMemoryAllocation ? PT_ERR_OUT_OF_MEMORY : PT_ERR_CUDA, #expr,      \
                            cudaGetErrorString(err__));                                                               \
    } while (0)

namespace pt
{

// device counters, one cache line each would be overkill: they are touched once per block
struct DeviceCounters
{
    unsigned long long raysClosest, raysShadow, samples, hits, boxClosest, triClosest, alphaClosest, boxShadow, triShadow,
        alphaShadow, texels, restarts;
};

// wavefront queue bookkeeping living in device memory
struct QueueCounts
{
    uint32_t active[2]; // double-buffered active-slot queue sizes
    uint32_t shadow;
    uint32_t pad;
};

// SoA path state, one entry per slot (slot <-> pixel of the current tile set)
struct PathState
{
    float4 *rayO;  // origin.xyz, maxRoughness
    float4 *rayD;  // direction.xyz, rng (bits)
    float4 *thr;   // throughput.xyz, bounce | flags (bits)
    float4 *rad;   // radiance.xyz, restarts (bits)
    float4 *diff0; // rxOrigin.xyz, rxDirection.x
    float4 *diff1; // rxDirection.yz, ryOrigin.xy
    float4 *diff2; // ryOrigin.z, ryDirection.xyz
    float4 *hit;   // tri (bits), t, b1, b2
    float4 *decal; // rgb, dist          (only when the scene has alpha-tested geometry)
    float *decalA; // alpha
    float4 *shO;   // shadow origin.xyz, tmax
    float4 *shD;   // shadow direction.xyz, -
    float4 *shC;   // contribution.xyz (throughput * DirectLight / pdf), -
    uint32_t *sample;    // next sample index (relative to first_sample) of the slot
    uint32_t *slotPixel; // y * width + x
    uint32_t *queue[2];  // active slots, double buffered
    uint32_t *shadowQueue;
};

struct Context
{
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    std::string lastError;

    // scene
    bool hasScene = false;
    DeviceScene scene = {};
    std::vector<void *> sceneAllocs; // everything cudaMalloc'ed for the scene
    std::vector<DevTexture> hostTextures;
    uint64_t texelArenaBytes = 0;
    uint64_t nodeCount = 0, bvhBytes = 0;
    float bvhBuildMs = 0, sceneUploadMs = 0;

    // target
    uint32_t width = 0, height = 0;
    float4 *accum = nullptr;
    PathState ps = {};
    std::vector<void *> targetAllocs;
    uint32_t slotCapacity = 0;
    // slot -> pixel map cache (rebuilt only when the tile list changes)
    std::vector<pt_tile> slotTiles;
    uint32_t slotCount = 0;
    bool slotMapValid = false;
    bool collectTraversalStats = false;
    bool kernelTiming = false;

    DeviceCounters *dCounters = nullptr;
    QueueCounts *dQueueCounts = nullptr;
    QueueCounts *hQueueCounts = nullptr; // pinned
    float *dLut = nullptr;

    // stats of the last call
    pt_stats stats = {};
    cudaEvent_t evStart = nullptr, evStop = nullptr;
};

pt_status fail(Context *ctx, pt_status code, const char *what, const char *detail);

// bvh_build.cu
pt_status uploadScene(Context *ctx, const pt_scene_desc *desc);
void freeScene(Context *ctx);
pt_status uploadTextureSlot(Context *ctx, uint32_t slot, const pt_texture_desc *tex);

// wavefront.cu
pt_status renderSamples(Context *ctx, const pt_render_params *params, uint32_t firstSample, uint32_t sampleCount,
                        const pt_tile *tiles, uint32_t tileCount);
pt_status firstHitAov(Context *ctx, const pt_render_params *params, uint32_t width, uint32_t height, pt_hit *out);
pt_status traceClosest(Context *ctx, const pt_ray *rays, uint64_t n, pt_hit *out);
pt_status traceOcclusion(Context *ctx, const pt_ray *rays, uint64_t n, uint8_t *out);

// unit_kernels.cu
pt_status testShading(Context *ctx, uint32_t mode, const float *input, float *output, uint32_t count);
uint32_t testInputStride(uint32_t mode);
uint32_t testOutputStride(uint32_t mode);

} // namespace pt

struct pt_context : pt::Context
{
};
