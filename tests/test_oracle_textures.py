"""The oracle's sampler definition: mip chain (floor(log2(max)) + 1 levels, linear-blit
down-sampling, Image.cpp:14-17,264-305), sRGB decode, repeat addressing, textureGrad LOD."""
import numpy as np

import conftest

sc = conftest.pkg.scene


def srgb_to_linear(c):
    c = c / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def test_default_textures(default_oracle):
    assert default_oracle.texture_info(0) == (1, 1, 1)
    assert (default_oracle.texture_level(1, 0) == [0x80, 0x80, 0xFF, 0xFF]).all()  # Q10: (128,128,255)/255
    n = default_oracle.texture_sample(1, [[0.3, 0.7, 0.1, 0, 0, 0.1]])
    assert np.allclose(n, [[128 / 255, 128 / 255, 1, 1]])
    assert np.allclose(default_oracle.texture_sample(4, [[0, 0, 0, 0, 0, 0]]), 0)


def test_mip_chain_shapes(default_oracle, default_scene):
    for i, tex in enumerate(default_scene.textures):
        w, h, levels = default_oracle.texture_info(sc.SCENE_TEXTURE_OFFSET + i)
        assert (h, w) == tex.pixels.shape[:2]
        assert levels == int(np.floor(np.log2(max(w, h)))) + 1
    # 2024 -> 1012 -> 506 -> 253 -> 126 ... (truncating halving, non-power-of-two)
    slot = sc.SCENE_TEXTURE_OFFSET + 3
    assert default_oracle.texture_level(slot, 3).shape[:2] == (253, 253)
    assert default_oracle.texture_level(slot, 4).shape[:2] == (126, 126)
    assert default_oracle.texture_level(slot, 10).shape[:2] == (1, 1)


def test_power_of_two_mip_is_box_filter_in_linear_space(default_oracle, default_scene):
    slot = sc.SCENE_TEXTURE_OFFSET + 1  # 512x512 sRGB colour texture
    l0 = default_oracle.texture_level(slot, 0).astype(np.float64)
    l1 = default_oracle.texture_level(slot, 1).astype(np.float64)
    assert (l0 == default_scene.textures[1].pixels).all()
    lin = srgb_to_linear(l0[..., :3])
    box = (lin[0::2, 0::2] + lin[1::2, 0::2] + lin[0::2, 1::2] + lin[1::2, 1::2]) / 4
    back = srgb_to_linear(l1[..., :3])
    # re-quantised to 8-bit sRGB: within one code of the exact box filter
    lo, hi = srgb_to_linear(np.maximum(l1[..., :3] - 1, 0)), srgb_to_linear(np.minimum(l1[..., :3] + 1, 255))
    assert ((box >= lo - 1e-9) & (box <= hi + 1e-9)).all() and np.abs(back - box).max() < 0.01
    a_box = (l0[0::2, 0::2, 3] + l0[1::2, 0::2, 3] + l0[0::2, 1::2, 3] + l0[1::2, 1::2, 3]) / 4
    assert np.abs(l1[..., 3] - a_box).max() <= 0.5 + 1e-9


def test_bilinear_and_repeat(default_oracle, default_scene):
    slot = sc.SCENE_TEXTURE_OFFSET + 1
    px = default_scene.textures[1].pixels
    W = px.shape[1]
    # texel centres reproduce the texel
    u, v = (17 + 0.5) / W, (33 + 0.5) / W
    got = default_oracle.texture_sample(slot, [[u, v, 0, 0, 0, 0]], use_grad=False)[0]
    want = np.concatenate([srgb_to_linear(px[33, 17, :3].astype(np.float64)), [px[33, 17, 3] / 255]])
    assert np.allclose(got, want, atol=1e-6)
    # repeat addressing: u and u + 3, v and v - 2 agree
    a = default_oracle.texture_sample(slot, [[0.123, 0.456, 0, 0, 0, 0]], use_grad=False)
    b = default_oracle.texture_sample(slot, [[3.123, -1.544, 0, 0, 0, 0]], use_grad=False)
    assert np.allclose(a, b, atol=1e-5)


def test_texture_grad_lod_selection(default_oracle):
    slot = sc.SCENE_TEXTURE_OFFSET + 1  # 512^2, 10 levels
    uv = [0.3, 0.6]
    mag = default_oracle.texture_sample(slot, [[*uv, 1e-4, 0, 0, 1e-4]])
    lod0 = default_oracle.texture_sample(slot, [[*uv, 0, 0, 0, 0]], use_grad=False)
    assert np.allclose(mag, lod0)  # rho < 1 texel: magnification
    coarse = default_oracle.texture_sample(slot, [[*uv, 10.0, 0, 0, 10.0]])
    top = default_oracle.texture_level(slot, 9).astype(np.float64)[0, 0]
    assert np.allclose(coarse[0, :3], srgb_to_linear(top[:3]), atol=1e-6)
    # NaN / Inf derivatives fall back to level 0 (derivative clamp in tracing.glsl lets NaN through)
    nan = default_oracle.texture_sample(slot, [[*uv, np.nan, 0, 0, 0]])
    assert np.allclose(nan, lod0)


def test_anisotropic_footprint_closed_form(default_scene, oracle_mod):
    """textureGrad with the reference's sampler state (anisotropy at the device maximum, Renderer.cpp:103-112), defined as
    the Vulkan specification's example implementation: eta = min(rho_max / rho_min, 16), N = ceil(eta) taps spread along
    the major axis at i / (N + 1) - 1/2 of its derivative, all at lambda = log2(rho_max / eta), averaged."""
    o = oracle_mod.OracleScene(default_scene)
    o.set_sampler(16)
    slot = sc.SCENE_TEXTURE_OFFSET + 1  # 512^2
    W = 512.0
    uv = np.array([0.37, 0.61])
    # ratio 4 at one texel of minor axis: lambda = log2(4 / 4) = 0 -> four bilinear taps of level 0 along x
    got = o.texture_sample(slot, [[*uv, 4 / W, 0, 0, 1 / W]])[0]
    taps = [[uv[0] + (i / 5 - 0.5) * 4 / W, uv[1], 0, 0, 0, 0] for i in (1, 2, 3, 4)]
    want = o.texture_sample(slot, taps, use_grad=False).astype(np.float64).mean(0)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
    # the major axis can be dPdy, and diagonal
    got = o.texture_sample(slot, [[*uv, 0.5 / W, 0.5 / W, -2 / W, 2 / W]])[0]  # rho_x = 0.707, rho_y = 2.83: eta = 4
    taps = [[uv[0] + (i / 5 - 0.5) * -2 / W, uv[1] + (i / 5 - 0.5) * 2 / W, 0, 0, 0, 0] for i in (1, 2, 3, 4)]
    # lambda = log2(2.83 / 4) < 0: level 0
    assert np.allclose(got, o.texture_sample(slot, taps, use_grad=False).astype(np.float64).mean(0), rtol=1e-6, atol=1e-7)
    # beyond the maximum: ratio 64 -> eta = 16, lambda = log2(64 / 16) = 2: sixteen bilinear taps of level 2
    got = o.texture_sample(slot, [[*uv, 64 / W, 0, 0, 1 / W]])[0]
    l2 = o.texture_level(slot, 2).astype(np.float64)
    lin = np.concatenate([srgb_to_linear(l2[..., :3]), l2[..., 3:] / 255], -1)

    def bilinear(img, u, v):
        h, w = img.shape[:2]
        x, y = (u % 1.0) * w - 0.5, (v % 1.0) * h - 0.5
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        fx, fy = x - x0, y - y0
        t = lambda xx, yy: img[yy % h, xx % w]
        return (t(x0, y0) * (1 - fx) + t(x0 + 1, y0) * fx) * (1 - fy) + (t(x0, y0 + 1) * (1 - fx) + t(x0 + 1, y0 + 1) * fx) * fy

    want = np.mean([bilinear(lin, uv[0] + (i / 17 - 0.5) * 64 / W, uv[1]) for i in range(1, 17)], 0)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    # isotropic footprints and maximum anisotropy 1 are the trilinear filter: one tap
    iso = o.texture_sample(slot, [[*uv, 3 / W, 0, 0, 3 / W]])[0]
    o.set_sampler(1)
    assert np.array_equal(o.texture_sample(slot, [[*uv, 3 / W, 0, 0, 3 / W]])[0], iso)
    blurry = o.texture_sample(slot, [[*uv, 64 / W, 0, 0, 1 / W]])[0]  # lambda = 6 now: far blurrier than the 16 taps of level 2
    top = o.texture_level(slot, 6).astype(np.float64)
    assert not np.allclose(blurry, got, atol=1e-3)
    assert np.abs(blurry[:3] - bilinear(np.concatenate([srgb_to_linear(top[..., :3]), top[..., 3:] / 255], -1), *uv)[:3]).max() < 1e-5


def test_textures_beyond_the_maximum_size_are_scaled_down(default_scene, oracle_mod):
    """TextureUploader::UploadTexture (TextureUploader.cpp:408-415): scale = max(ceil(w / max), ceil(h / max)), extent =
    (w / scale, h / scale), a linear blit of the full image; then the usual mip chain.  MaxTextureDataSize is 4096 in the
    reference; the limit is lowered here so that the 1024 / 512 / 2024-pixel Default-scene textures exercise it."""
    full = oracle_mod.OracleScene(default_scene)
    small = oracle_mod.OracleScene(default_scene, max_texture_size=512)
    base = sc.SCENE_TEXTURE_OFFSET
    assert full.texture_info(base + 0)[:2] == (1024, 1024) and small.texture_info(base + 0) == (512, 512, 10)  # scale 2
    assert small.texture_info(base + 1) == full.texture_info(base + 1)                                         # fits
    assert full.texture_info(base + 3)[:2] == (2024, 2024) and small.texture_info(base + 3) == (506, 506, 9)   # scale 4
    # scale 2 of an even extent = the 2 x 2 box filter = level 1 of the full-size chain, and so on down the chain
    for level in range(3):
        assert np.array_equal(small.texture_level(base + 0, level), full.texture_level(base + 0, level + 1))
    # the budget: 4 scene textures sharing 4 MB -> 1 MB each -> largest extent whose full RGBA8 chain fits: 256
    tight = oracle_mod.OracleScene(default_scene, texture_budget_mb=4)
    assert tight.texture_info(base + 0)[:2] == (256, 256) and tight.texture_info(base + 1)[:2] == (256, 256)
    assert tight.texture_info(base + 3)[:2] == (253, 253)  # 2024 / ceil(2024 / 256 = 7.9) = 2024 / 8
