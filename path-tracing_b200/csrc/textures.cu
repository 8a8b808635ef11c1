// textures.cu — texture upload: mip chains of RGBA8 / float images and the decode of block-compressed
// (BC1 / BC3 / BC5) images into the RGBA8 texels the software sampler reads.
//
// Replaces TextureUploader's image creation + mip generation (PT/Renderer/TextureUploader.cpp:402-527,
// PT/Renderer/Image.cpp:14-17, 179-182, 264-305: floor(log2(max(w, h))) + 1 levels, each blitted from the
// previous one with a linear filter) and the sampler hardware's BC decode (TextureUploader.cpp:586-591).
#include "core_internal.h"

#include <algorithm>
#include <vector>

namespace pt
{

namespace
{

// ---------------------------------------------------------------------------------------------
// textures
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mipTexel(const DevTexture &t, const float *lut, uint32_t level, int x, int y, uint32_t lw)
{
    const size_t idx = (size_t)t.levelOffset[level] + (size_t)y * lw + x;
    if (t.flags & PT_TEX_FLAG_FLOAT)
        return reinterpret_cast<const float4 *>(t.base)[idx];
    const uchar4 c = reinterpret_cast<const uchar4 *>(t.base)[idx];
    const float *l = lut + ((t.flags & PT_TEX_FLAG_SRGB) ? 256 : 0);
    return make_float4(l[c.x], l[c.y], l[c.z], lut[c.w]);
}
__device__ __forceinline__ float lerpExact(float a, float b, float f)
{
    // a * (1 - f) + b * f without contraction, to match the oracle bit for bit
    return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, f)), __fmul_rn(b, f));
}
__device__ __forceinline__ uint8_t encodeUnorm8(float v)
{
    if (!(v > 0.0f))
        return 0;
    if (v >= 1.0f)
        return 255;
    return (uint8_t)__fadd_rn(__fmul_rn(v, 255.0f), 0.5f);
}
__device__ __forceinline__ uint8_t encodeSrgb8(const float *lut, float v)
{
    // nearest sRGB code in linear space: smallest i with v < (srgb[i] + srgb[i+1]) / 2
    int lo = 0, hi = 255;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        const float m = __fmul_rn(0.5f, __fadd_rn(lut[256 + mid], lut[256 + mid + 1]));
        if (v < m)
            hi = mid;
        else
            lo = mid + 1;
    }
    return (uint8_t)lo;
}

// Linear vkCmdBlitImage of one whole level into another (PT/Renderer/Image.cpp:264-305): level k from level k - 1 of
// the same texture (mip generation), or a full-size staging image into the down-scaled level 0
// (TextureUploader.cpp:408-415, 470-520: textures beyond the maximum texture size).
__global__ void k_blit(DevTexture src, uint32_t srcLevel, DevTexture dst, uint32_t dstLevel, const float *__restrict__ lut)
{
    const uint32_t sw = max(1u, src.width >> srcLevel), sh = max(1u, src.height >> srcLevel);
    const uint32_t dw = max(1u, dst.width >> dstLevel), dh = max(1u, dst.height >> dstLevel);
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh)
        return;
    const float sxScale = __fdiv_rn((float)sw, (float)dw), syScale = __fdiv_rn((float)sh, (float)dh);
    const float sx = __fsub_rn(__fmul_rn(__fadd_rn((float)x, 0.5f), sxScale), 0.5f);
    const float sy = __fsub_rn(__fmul_rn(__fadd_rn((float)y, 0.5f), syScale), 0.5f);
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const float fx = __fsub_rn(sx, fx0), fy = __fsub_rn(sy, fy0);
    const int x0 = min(max((int)fx0, 0), (int)sw - 1), x1 = min(max((int)fx0 + 1, 0), (int)sw - 1);
    const int y0 = min(max((int)fy0, 0), (int)sh - 1), y1 = min(max((int)fy0 + 1, 0), (int)sh - 1);
    const float4 t00 = mipTexel(src, lut, srcLevel, x0, y0, sw), t10 = mipTexel(src, lut, srcLevel, x1, y0, sw);
    const float4 t01 = mipTexel(src, lut, srcLevel, x0, y1, sw), t11 = mipTexel(src, lut, srcLevel, x1, y1, sw);
    float4 v;
    v.x = lerpExact(lerpExact(t00.x, t10.x, fx), lerpExact(t01.x, t11.x, fx), fy);
    v.y = lerpExact(lerpExact(t00.y, t10.y, fx), lerpExact(t01.y, t11.y, fx), fy);
    v.z = lerpExact(lerpExact(t00.z, t10.z, fx), lerpExact(t01.z, t11.z, fx), fy);
    v.w = lerpExact(lerpExact(t00.w, t10.w, fx), lerpExact(t01.w, t11.w, fx), fy);
    const size_t idx = (size_t)dst.levelOffset[dstLevel] + (size_t)y * dw + x;
    if (dst.flags & PT_TEX_FLAG_FLOAT)
    {
        reinterpret_cast<float4 *>(dst.base)[idx] = v;
        return;
    }
    uchar4 o;
    if (dst.flags & PT_TEX_FLAG_SRGB)
        o = make_uchar4(encodeSrgb8(lut, v.x), encodeSrgb8(lut, v.y), encodeSrgb8(lut, v.z), encodeUnorm8(v.w));
    else
        o = make_uchar4(encodeUnorm8(v.x), encodeUnorm8(v.y), encodeUnorm8(v.z), encodeUnorm8(v.w));
    reinterpret_cast<uchar4 *>(dst.base)[idx] = o;
}

// ---------------------------------------------------------------------------------------------
// Block-compressed textures (TextureFormat::BC1 / BC3 / BC5, PT/Scene.h:35-42; the reference hands
// the blocks to the sampler hardware, VK_FORMAT_BC1_RGBA / BC3 / BC5, TextureUploader.cpp:586-591).
// Decoded here once, at upload, into the RGBA8 texels every other texture uses: one thread per 4x4
// block.  Interpolated palette entries are the exact rationals of the format definition rounded to
// the nearest 8-bit value (halves up).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bcColorPalette(const uint8_t *b, bool allowPunchThrough, uchar4 pal[4])
{
    const uint32_t c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
    auto expand = [](uint32_t c) {
        const uint32_t r = c >> 11, g = (c >> 5) & 63u, bl = c & 31u;
        return make_uchar4((uint8_t)((r << 3) | (r >> 2)), (uint8_t)((g << 2) | (g >> 4)), (uint8_t)((bl << 3) | (bl >> 2)), 255);
    };
    pal[0] = expand(c0);
    pal[1] = expand(c1);
    if (c0 > c1 || !allowPunchThrough)
    {
        pal[2] = make_uchar4((uint8_t)((2 * pal[0].x + pal[1].x + 1) / 3), (uint8_t)((2 * pal[0].y + pal[1].y + 1) / 3),
                             (uint8_t)((2 * pal[0].z + pal[1].z + 1) / 3), 255);
        pal[3] = make_uchar4((uint8_t)((pal[0].x + 2 * pal[1].x + 1) / 3), (uint8_t)((pal[0].y + 2 * pal[1].y + 1) / 3),
                             (uint8_t)((pal[0].z + 2 * pal[1].z + 1) / 3), 255);
    }
    else
    {
        pal[2] = make_uchar4((uint8_t)((pal[0].x + pal[1].x + 1) / 2), (uint8_t)((pal[0].y + pal[1].y + 1) / 2),
                             (uint8_t)((pal[0].z + pal[1].z + 1) / 2), 255);
        pal[3] = make_uchar4(0, 0, 0, 0); // BC1_RGBA: transparent black
    }
}

// the 8-byte single-channel block of BC3's alpha and BC5's two channels
__device__ __forceinline__ void bcAlphaPalette(const uint8_t *b, uint8_t pal[8])
{
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
}

__global__ void k_bc_decode(const uint8_t *__restrict__ blocks, uint32_t format, uint32_t w, uint32_t h, uchar4 *__restrict__ out)
{
    const uint32_t bw = (w + 3) / 4, bh = (h + 3) / 4;
    const uint32_t bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= bw * bh)
        return;
    const uint32_t bx = bi % bw, by = bi / bw;
    const uint8_t *b = blocks + (size_t)bi * (format == PT_TEXTURE_BC1 ? 8 : 16);
    uchar4 texel[16];
    if (format == PT_TEXTURE_BC5)
    {
        uint8_t pr[8], pg[8];
        bcAlphaPalette(b, pr);
        bcAlphaPalette(b + 8, pg);
        unsigned long long ir = 0, ig = 0;
        for (int k = 0; k < 6; k++)
        {
            ir |= (unsigned long long)b[2 + k] << (8 * k);
            ig |= (unsigned long long)b[10 + k] << (8 * k);
        }
        for (int t = 0; t < 16; t++)
            texel[t] = make_uchar4(pr[(ir >> (3 * t)) & 7u], pg[(ig >> (3 * t)) & 7u], 0, 255);
    }
    else
    {
        const uint8_t *color = format == PT_TEXTURE_BC3 ? b + 8 : b;
        uchar4 pal[4];
        bcColorPalette(color, format == PT_TEXTURE_BC1, pal);
        const uint32_t idx = color[4] | (color[5] << 8) | (color[6] << 16) | ((uint32_t)color[7] << 24);
        for (int t = 0; t < 16; t++)
            texel[t] = pal[(idx >> (2 * t)) & 3u];
        if (format == PT_TEXTURE_BC3)
        {
            uint8_t pa[8];
            bcAlphaPalette(b, pa);
            unsigned long long ia = 0;
            for (int k = 0; k < 6; k++)
                ia |= (unsigned long long)b[2 + k] << (8 * k);
            for (int t = 0; t < 16; t++)
                texel[t].w = pa[(ia >> (3 * t)) & 7u];
        }
    }
    for (int t = 0; t < 16; t++)
    {
        const uint32_t x = bx * 4 + (t & 3), y = by * 4 + (t >> 2);
        if (x < w && y < h)
            out[(size_t)y * w + x] = texel[t];
    }
}

} // namespace

// TextureUploader::DetermineMaxTextureSizes (TextureUploader.cpp:551-569): 4096 (MaxTextureDataSize), halved while the full
// mip chain of a texture of that extent exceeds the per-texture share of the budget (0 = ForceFullTextureSize)
uint32_t maxTextureExtent(Context *ctx, uint32_t bytesPerTexel, bool scalable)
{
    uint32_t extent = ctx->maxTextureSize;
    if (!scalable || ctx->textureBudgetBytes == 0)
        return extent;
    const uint64_t perTexture = ctx->textureBudgetBytes / std::max<uint64_t>(1, ctx->textureBudgetCount);
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

pt_status createTexture(Context *ctx, const pt_texture_desc &d, DevTexture &out, void **outAlloc, bool scalable)
{
    if (d.width == 0 || d.height == 0 || !d.pixels)
        return fail(ctx, PT_ERR_INVALID_ARGUMENT, "texture", "empty texture");
    if (d.format > PT_TEXTURE_BC5)
        return fail(ctx, PT_ERR_UNSUPPORTED, "texture", "unknown texture format");
    if (d.format >= PT_TEXTURE_BC1)
    {
        // stored mip chain, decoded level by level; no mips are generated (TextureUploader.cpp:420-456)
        DevTexture t = {};
        t.width = d.width;
        t.height = d.height;
        t.flags = (d.format == PT_TEXTURE_BC3 || (d.format == PT_TEXTURE_BC1 && d.srgb)) ? PT_TEX_FLAG_SRGB : 0u;
        uint32_t full = 1;
        for (uint32_t m = std::max(d.width, d.height); m > 1; m >>= 1)
            full++;
        const uint32_t levels = std::max(1u, d.levels);
        if (levels > full || levels > PT_MAX_TEX_LEVELS)
            return fail(ctx, PT_ERR_INVALID_ARGUMENT, "texture", "more mip levels than the extent allows");
        t.levels = levels;
        const uint64_t blockBytes = d.format == PT_TEXTURE_BC1 ? 8 : 16;
        uint64_t texels = 0, bytes = 0;
        std::vector<uint64_t> blockOffset(levels);
        for (uint32_t l = 0; l < levels; l++)
        {
            const uint32_t lw = std::max(1u, d.width >> l), lh = std::max(1u, d.height >> l);
            t.levelOffset[l] = (uint32_t)texels;
            texels += (uint64_t)lw * lh;
            blockOffset[l] = bytes;
            bytes += (uint64_t)((lw + 3) / 4) * ((lh + 3) / 4) * blockBytes;
        }
        void *mem = nullptr;
        uint8_t *dBlocks = nullptr;
        PT_CUDA_CHECK(ctx, cudaMalloc(&mem, texels * 4));
        cudaError_t err = cudaMalloc((void **)&dBlocks, bytes);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(dBlocks, d.pixels, bytes, cudaMemcpyHostToDevice, ctx->stream);
        for (uint32_t l = 0; l < levels && err == cudaSuccess; l++)
        {
            const uint32_t lw = std::max(1u, d.width >> l), lh = std::max(1u, d.height >> l);
            const uint32_t nBlocks = ((lw + 3) / 4) * ((lh + 3) / 4);
            k_bc_decode<<<(nBlocks + 127) / 128, 128, 0, ctx->stream>>>(dBlocks + blockOffset[l], d.format, lw, lh,
                                                                        reinterpret_cast<uchar4 *>(mem) + t.levelOffset[l]);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess)
            err = cudaStreamSynchronize(ctx->stream);
        cudaFree(dBlocks);
        if (err != cudaSuccess)
        {
            cudaFree(mem);
            PT_CUDA_CHECK(ctx, err);
        }
        t.base = (uint64_t)mem;
        out = t;
        *outAlloc = mem;
        return PT_OK;
    }
    DevTexture t = {};
    t.flags = d.format == PT_TEXTURE_RGBAF32 ? PT_TEX_FLAG_FLOAT : (d.srgb ? PT_TEX_FLAG_SRGB : 0u);
    // TextureUploader::UploadTexture (TextureUploader.cpp:408-415): textures beyond the maximum extent are scaled down by
    // an integer factor with a linear blit before their mips are generated
    const uint32_t maxExtent = maxTextureExtent(ctx, (t.flags & PT_TEX_FLAG_FLOAT) ? 16 : 4, scalable);
    const uint32_t scale = std::max((d.width + maxExtent - 1) / maxExtent, (d.height + maxExtent - 1) / maxExtent);
    if (scale > 1 && (t.flags & PT_TEX_FLAG_FLOAT))
        return fail(ctx, PT_ERR_UNSUPPORTED, "texture", "a float texture exceeds the maximum texture size and its format cannot be "
                                                        "scaled (TextureUploader.cpp:466-480: the reference rejects it too)");
    t.width = std::max(d.width / scale, 1u);
    t.height = std::max(d.height / scale, 1u);
    // floor(log2(max(w, h))) + 1 levels (PT/Renderer/Image.cpp:14-17)
    uint32_t levels = 1;
    for (uint32_t m = std::max(t.width, t.height); m > 1; m >>= 1)
        levels++;
    if (levels > PT_MAX_TEX_LEVELS)
        return fail(ctx, PT_ERR_UNSUPPORTED, "texture", "texture larger than 32768 texels per side");
    t.levels = levels;
    uint64_t texels = 0;
    for (uint32_t l = 0; l < levels; l++)
    {
        t.levelOffset[l] = (uint32_t)texels;
        texels += (uint64_t)std::max(1u, t.width >> l) * std::max(1u, t.height >> l);
    }
    const uint64_t bpp = (t.flags & PT_TEX_FLAG_FLOAT) ? 16 : 4;
    void *mem = nullptr;
    PT_CUDA_CHECK(ctx, cudaMalloc(&mem, texels * bpp));
    t.base = (uint64_t)mem;
    void *staging = nullptr;
    if (scale > 1)
    {
        DevTexture full = {};
        full.width = d.width, full.height = d.height, full.levels = 1, full.flags = t.flags;
        cudaError_t err = cudaMalloc(&staging, (uint64_t)d.width * d.height * bpp);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(staging, d.pixels, (uint64_t)d.width * d.height * bpp, cudaMemcpyHostToDevice, ctx->stream);
        if (err != cudaSuccess)
        {
            cudaFree(staging);
            cudaFree(mem);
            PT_CUDA_CHECK(ctx, err);
        }
        full.base = (uint64_t)staging;
        const dim3 block(16, 16), grid((t.width + 15) / 16, (t.height + 15) / 16);
        k_blit<<<grid, block, 0, ctx->stream>>>(full, 0, t, 0, ctx->dLut);
    }
    else
        PT_CUDA_CHECK(ctx, cudaMemcpyAsync(mem, d.pixels, (uint64_t)d.width * d.height * bpp, cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t l = 1; l < levels; l++)
    {
        const uint32_t dw = std::max(1u, t.width >> l), dh = std::max(1u, t.height >> l);
        const dim3 block(16, 16), grid((dw + 15) / 16, (dh + 15) / 16);
        k_blit<<<grid, block, 0, ctx->stream>>>(t, l - 1, t, l, ctx->dLut);
    }
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    {
        cudaFree(staging);
        cudaFree(mem);
        return fail(ctx, PT_ERR_CUDA, "texture", "mip generation failed");
    }
    cudaFree(staging);
    // the host pixels may be released as soon as we return
    PT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    out = t;
    *outAlloc = mem;
    return PT_OK;
}

pt_texture_desc defaultTexture(const uint32_t *rgba, bool srgb)
{
    pt_texture_desc d = {};
    d.width = d.height = 1;
    d.format = PT_TEXTURE_RGBA8;
    d.srgb = srgb;
    d.pixels = rgba;
    return d;
}

} // namespace pt
