// pt_oracle_post.cpp — CPU restatement of the reference's post-processing + output chain.
//
// TEST INFRASTRUCTURE, NOT PRODUCT (see pt_oracle.h).
//
// Follows, dispatch by dispatch, Renderer::RecordPostProcessCommands
// (PT/Renderer/Renderer.cpp:928-1060) and Renderer::RecordSaveOutputCommands (:1205-1250):
//   postprocess.comp -> bloomDownsample.comp x (maxMip - 1) -> bloomUpsample.comp x (maxMip - 1)
//   -> composition.comp -> blit into OutputSaver's RGBA16F "linear output image"
//   (PT/Renderer/OutputSaver.cpp:66-83) -> toneMapping.comp in place -> blit to the output
//   format (OutputSaver.cpp:113-140, 255-273).
// Every storage image of the chain is RGBA16F, so every imageStore rounds to half precision
// (round to nearest even); the bloom sampler is linear / clamp-to-edge / no mips
// (Renderer.cpp:114-119).
//
// PINNED to the reference's own compute shaders: postprocess.comp, bloomDownsample.comp, bloomUpsample.comp,
// composition.comp and toneMapping.comp compiled as C++ (oracle/ref_overlay/glsl2cpp.py --compute ->
// oracle/_ref/libglsl_comp_ref.so) give bit-identical images after composition and after tone mapping
// (tests/test_oracle_vs_glsl_compute.py, golden vectors in tests/golden/glsl_compute_vectors.npz).  What stays
// PARITY UNPINNED are the steps the reference leaves to the Vulkan implementation — the sampler's bilinear weights,
// the binary16 rounding of the image stores, the blit's sRGB encode — restated here from the Vulkan specification's
// formulas in fp32 / fp64.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pt_oracle.h"

namespace
{

// IEEE binary16 <-> binary32, round to nearest even
uint16_t floatToHalf(float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) // Inf / NaN
        return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));
    if (x >= 0x477ff000u) // >= 65520 rounds to infinity
        return (uint16_t)(sign | 0x7c00u);
    if (x < 0x33000001u) // <= 2^-25 rounds to zero
        return (uint16_t)sign;
    uint32_t exp = x >> 23, man = x & 0x7fffffu;
    if (exp < 113) // subnormal half
    {
        man |= 0x800000u;
        const uint32_t shift = 126 - exp; // 14 .. 24
        const uint32_t half = man >> shift, rem = man & ((1u << shift) - 1u), mid = 1u << (shift - 1);
        return (uint16_t)(sign | (half + ((rem > mid || (rem == mid && (half & 1u))) ? 1u : 0u)));
    }
    const uint32_t half = ((exp - 112) << 10) | (man >> 13), rem = man & 0x1fffu;
    return (uint16_t)(sign | (half + ((rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ? 1u : 0u)));
}

float halfToFloat(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, x;
    if (exp == 0)
    {
        if (man == 0)
            x = sign;
        else
        {
            exp = 113;
            while (!(man & 0x400u))
            {
                man <<= 1;
                exp--;
            }
            x = sign | (exp << 23) | ((man & 0x3ffu) << 13);
        }
    }
    else if (exp == 31)
        x = sign | 0x7f800000u | (man << 13);
    else
        x = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

float roundHalf(float f) { return halfToFloat(floatToHalf(f)); }

struct Rgb
{
    float r, g, b;
};
Rgb operator+(Rgb a, Rgb b) { return { a.r + b.r, a.g + b.g, a.b + b.b }; }
Rgb operator*(Rgb a, float s) { return { a.r * s, a.g * s, a.b * s }; }

// one mip level of an RGBA16F image; values are kept as the floats the halves decode to
struct Level
{
    uint32_t w = 0, h = 0;
    std::vector<Rgb> px;
    void store(uint32_t x, uint32_t y, Rgb c) { px[(size_t)y * w + x] = { roundHalf(c.r), roundHalf(c.g), roundHalf(c.b) }; }
    Rgb load(int x, int y) const { return px[(size_t)y * w + x]; }
    // texture(sampler2D, uv): linear filter, clamp to edge, unnormalised coordinate u * size - 0.5
    Rgb sample(float u, float v) const
    {
        const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
        const float fx0 = std::floor(x), fy0 = std::floor(y);
        const float fx = x - fx0, fy = y - fy0;
        const int x0 = std::clamp((int)fx0, 0, (int)w - 1), x1 = std::clamp((int)fx0 + 1, 0, (int)w - 1);
        const int y0 = std::clamp((int)fy0, 0, (int)h - 1), y1 = std::clamp((int)fy0 + 1, 0, (int)h - 1);
        const Rgb top = load(x0, y0) * (1.0f - fx) + load(x1, y0) * fx;
        const Rgb bot = load(x0, y1) * (1.0f - fx) + load(x1, y1) * fx;
        return top * (1.0f - fy) + bot * fy;
    }
};

uint32_t mipLevels(uint32_t w, uint32_t h) // PT/Renderer/Image.cpp:14-17
{
    uint32_t levels = 1;
    for (uint32_t m = std::max(w, h); m > 1; m >>= 1)
        levels++;
    return levels;
}

// VK_FORMAT_R8G8B8A8_SRGB store of a linear value (Vulkan spec, "sRGB EOTF^-1" + UNORM conversion)
uint8_t encodeSrgb8(float c)
{
    if (!(c > 0.0f))
        return 0;
    if (c >= 1.0f)
        return 255;
    const double l = c;
    const double e = l <= 0.0031308 ? 12.92 * l : 1.055 * std::pow(l, 1.0 / 2.4) - 0.055;
    return (uint8_t)std::lrint(e * 255.0); // round half to even, like the device's __double2int_rn
}

} // namespace

extern "C" int32_t pto_postprocess(const float *accum, uint32_t width, uint32_t height, const pt_postprocess_params *p,
                                   uint32_t total_samples, uint32_t output_format, void *out_pixels)
{
    if (!accum || !p || !out_pixels || width == 0 || height == 0)
        return PT_ERR_INVALID_ARGUMENT;
    const uint32_t levels = mipLevels(width, height);
    Level pp;
    pp.w = width, pp.h = height;
    pp.px.resize((size_t)width * height);
    std::vector<Level> bloom(levels);
    for (uint32_t l = 0; l < levels; l++)
    {
        bloom[l].w = std::max(1u, width >> l);
        bloom[l].h = std::max(1u, height >> l);
        bloom[l].px.assign((size_t)bloom[l].w * bloom[l].h, Rgb { 0, 0, 0 });
    }

    // ---- postprocess.comp:16-40 ------------------------------------------------------------------
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++)
        {
            const float *a = accum + 4 * ((size_t)y * width + x);
            const float ts = (float)total_samples;
            Rgb color = { a[0] / ts * p->exposure, a[1] / ts * p->exposure, a[2] / ts * p->exposure };
            if (std::isnan(color.r) || std::isnan(color.g) || std::isnan(color.b))
                color = { 5000.0f, 0.0f, 0.0f };
            if (std::isinf(color.r) || std::isinf(color.g) || std::isinf(color.b))
                color = { 0.0f, 5000.0f, 0.0f };
            const float knee = 0.5f, threshold = p->bloom_threshold;
            const float br = std::max(color.r, std::max(color.g, color.b));
            const float cx = threshold - knee, cy = knee * 2.0f, cz = 0.25f / knee;
            float rq = std::min(std::max(br - cx, 0.0f), cy);
            rq = cz * rq * rq;
            const Rgb bloomColor = color * (std::max(rq, br - threshold) / std::max(br, 0.0001f));
            pp.store(x, y, color);
            bloom[0].store(x, y, bloomColor);
        }

    // ---- bloom chain (Renderer.cpp:955-1039) --------------------------------------------------------
    // The reference computes min(levels - 3, 12) in unsigned arithmetic and would index
    // non-existent levels for frames smaller than 8 pixels; such frames get no bloom passes here.
    const uint32_t maxMip = levels > 3 ? std::min(levels - 3, 12u) : 1u;
    for (uint32_t i = 0; i + 1 < maxMip; i++) // bloomDownsample.comp:18-58
    {
        const Level &src = bloom[i];
        Level &dst = bloom[i + 1];
        const float tx = 1.0f / (float)src.w, ty = 1.0f / (float)src.h;
        for (uint32_t y = 0; y < dst.h; y++)
            for (uint32_t x = 0; x < dst.w; x++)
            {
                const float u = ((float)x + 0.5f) / (float)dst.w, v = ((float)y + 0.5f) / (float)dst.h;
                auto tap = [&](float ox, float oy) { return src.sample(u + ox * tx, v + oy * ty); };
                const Rgb a = tap(-2, 2), b = tap(0, 2), c = tap(2, 2);
                const Rgb d = tap(-2, 0), e = tap(0, 0), f = tap(2, 0);
                const Rgb g = tap(-2, -2), h = tap(0, -2), i9 = tap(2, -2);
                const Rgb j = tap(-1, 1), k = tap(1, 1), l = tap(-1, -1), m = tap(1, -1);
                Rgb down = e * 0.125f;
                down = down + (a + c + g + i9) * 0.03125f;
                down = down + (b + d + f + h) * 0.0625f;
                down = down + (j + k + l + m) * 0.125f;
                dst.store(x, y, down);
            }
    }
    for (uint32_t i = maxMip - 1; i > 0; i--) // bloomUpsample.comp:18-52
    {
        const Level &src = bloom[i];
        Level &dst = bloom[i - 1];
        const float tx = 1.0f / (float)src.w, ty = 1.0f / (float)src.h;
        for (uint32_t y = 0; y < dst.h; y++)
            for (uint32_t x = 0; x < dst.w; x++)
            {
                const float u = ((float)x + 0.5f) / (float)dst.w, v = ((float)y + 0.5f) / (float)dst.h;
                auto tap = [&](float ox, float oy) { return src.sample(u + ox * tx, v + oy * ty); };
                const Rgb a = tap(-1, 1), b = tap(0, 1), c = tap(1, 1);
                const Rgb d = tap(-1, 0), e = tap(0, 0), f = tap(1, 0);
                const Rgb g = tap(-1, -1), h = tap(0, -1), i9 = tap(1, -1);
                Rgb up = e * 4.0f;
                up = up + (b + d + f + h) * 2.0f;
                up = up + (a + c + g + i9);
                up = up * (1.0f / 16.0f);
                dst.store(x, y, dst.load((int)x, (int)y) + up);
            }
    }

    // ---- composition.comp:15-25, toneMapping.comp:13-24, output blit ---------------------------------
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++)
        {
            const size_t idx = (size_t)y * width + x;
            const Rgb c0 = pp.px[idx], bl = bloom[0].px[idx];
            const float k = p->bloom_intensity * 0.1f;
            Rgb c = { roundHalf(k * bl.r + 1.0f * c0.r), roundHalf(k * bl.g + 1.0f * c0.g), roundHalf(k * bl.b + 1.0f * c0.b) };
            if (p->tone_mapping == PT_TONE_MAPPING_SDR)
            {
                // 1 - exp(-c), evaluated in fp64 and rounded once to fp32 (then to the RGBA16F store)
                c.r = roundHalf((float)(1.0 - std::exp(-(double)c.r)));
                c.g = roundHalf((float)(1.0 - std::exp(-(double)c.g)));
                c.b = roundHalf((float)(1.0 - std::exp(-(double)c.b)));
            }
            if (output_format == PT_OUTPUT_RGBA8_SRGB)
            {
                uint8_t *o = (uint8_t *)out_pixels + 4 * idx;
                o[0] = encodeSrgb8(c.r), o[1] = encodeSrgb8(c.g), o[2] = encodeSrgb8(c.b), o[3] = 255;
            }
            else
            {
                float *o = (float *)out_pixels + 4 * idx;
                o[0] = c.r, o[1] = c.g, o[2] = c.b, o[3] = 1.0f;
            }
        }
    return PT_OK;
}

// half-precision round trip of n floats (KAT hook for the converters above)
extern "C" int32_t pto_round_half(const float *in, float *out, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++)
        out[i] = roundHalf(in[i]);
    return PT_OK;
}
