"""The software sampler of the CUDA core against the oracle's, record by record (pt_test_texture vs
pto_texture_sample): bilinear + trilinear filtering, repeat wrap, sRGB / UNORM / float texels, non-power-of-two mip
chains, block-compressed textures with stored mips."""
import copy
import importlib

import numpy as np
import pytest

import bc_ref
import conftest

pytestmark = pytest.mark.gpu
core = conftest.core
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def _records(rs, n):
    uv = rs.uniform(-2.5, 3.5, (n, 2))
    scale = 10.0 ** rs.uniform(-5, 0.5, (n, 1))  # footprints from far below a texel to the whole image
    d = rs.normal(size=(n, 4)) * scale
    rec = np.concatenate([uv, d], 1).astype(np.float32)
    rec[:16, 2:] = 0.0  # zero derivatives: level 0
    rec[16:24, :2] = [[0, 0], [1, 1], [0.5, 0.5], [-1, 2], [1e-7, 1 - 1e-7], [0.25, 0.75], [3, -3], [0.999999, 0.000001]]
    return rec


def _check(r, ora, slot, rec):
    for use_grad in (True, False):
        got = r.texture_sample(slot, rec, use_grad)
        want = ora.texture_sample(slot, rec, use_grad)
        # same arithmetic on both sides (unfused lerps, the shared 8-bit -> linear table): bit for bit except for
        # log2f, which differs in its last place between libm and CUDA and is only used to pick the mip blend
        close = np.isclose(got, want, rtol=2e-6, atol=1e-7)
        assert close.all(), (slot, use_grad, np.abs(got - want).max())
        assert (got == want).mean() > 0.97


def test_default_scene_textures(default_renderer, default_oracle, default_scene):
    """The reference's own Default scene: 1024^2, 512^2 and the non-power-of-two 2024^2 sRGB textures + the 1x1 built-ins."""
    rec = _records(np.random.default_rng(3), 4000)
    for slot in list(range(sc.SCENE_TEXTURE_OFFSET)) + [sc.SCENE_TEXTURE_OFFSET + i for i in range(len(default_scene.textures))]:
        _check(default_renderer, default_oracle, slot, rec)


def test_float_unorm_and_bc_textures(oracle_mod):
    s = scenes.feature_scene(texture_size=64)
    s = copy.copy(s)
    rs = np.random.default_rng(9)
    hdr = sc.Texture((rs.random((24, 40, 4)) * 4).astype(np.float32))                      # float, 40x24
    odd = sc.Texture(rs.integers(0, 256, (37, 19, 4), dtype=np.uint8), srgb=False)           # UNORM, 19x37
    px = np.ascontiguousarray(s.textures[0].pixels)
    bc = [sc.Texture(bc_ref.encode_chain(f, px, 4), srgb=(f != bc_ref.BC5), bc_format=f, bc_extent=(64, 64), levels=4)
          for f in (bc_ref.BC1, bc_ref.BC3, bc_ref.BC5)]
    s.textures = list(s.textures) + [hdr, odd] + bc
    ora = oracle_mod.OracleScene(s)
    rec = _records(rs, 3000)
    with core.Renderer(0) as r:
        r.update_scene_data(s)
        for slot in range(sc.SCENE_TEXTURE_OFFSET, sc.SCENE_TEXTURE_OFFSET + len(s.textures)):
            _check(r, ora, slot, rec)
        with pytest.raises(core.PtError):
            r.texture_sample(sc.SCENE_TEXTURE_OFFSET + len(s.textures), rec)


def test_anisotropic_sampler_flag(default_scene, oracle_mod):
    """pt_set_sampler(16): the reference's sampler state (anisotropy at the device maximum, Renderer.cpp:103-112) as the
    Vulkan specification's example implementation — record by record against the oracle's (closed forms:
    tests/test_oracle_textures.py), and a whole image of the textured Default scene."""
    import metrics

    rs = np.random.default_rng(21)
    rec = _records(rs, 4000)
    # elongated footprints: ratio 1 .. 100 between the two derivative axes
    n = 2000
    ang = rs.uniform(0, 2 * np.pi, n)
    long_ = 10.0 ** rs.uniform(-3.5, -0.5, n)
    ratio = 10.0 ** rs.uniform(0, 2, n)
    ax = np.stack([np.cos(ang), np.sin(ang)], 1)
    perp = np.stack([-np.sin(ang), np.cos(ang)], 1)
    swap = rs.uniform(size=n) < 0.5
    dx = np.where(swap[:, None], perp * (long_ / ratio)[:, None], ax * long_[:, None])
    dy = np.where(swap[:, None], ax * long_[:, None], perp * (long_ / ratio)[:, None])
    rec2 = np.concatenate([rs.uniform(-1, 2, (n, 2)), dx, dy], 1).astype(np.float32)
    ora = oracle_mod.OracleScene(default_scene)
    ora.set_sampler(16)
    with core.Renderer(0) as r:
        r.update_scene_data(default_scene)
        r.set_sampler(16)
        for slot in [1, sc.SCENE_TEXTURE_OFFSET + 1, sc.SCENE_TEXTURE_OFFSET + 3]:
            _check(r, ora, slot, rec)
            _check(r, ora, slot, rec2)
        # the flag changes the fetches (and only the textureGrad ones)
        iso = oracle_mod.OracleScene(default_scene)
        a, b = ora.texture_sample(sc.SCENE_TEXTURE_OFFSET + 1, rec2), iso.texture_sample(sc.SCENE_TEXTURE_OFFSET + 1, rec2)
        assert (np.abs(a - b).max(1) > 1e-3).mean() > 0.1
        assert np.array_equal(ora.texture_sample(sc.SCENE_TEXTURE_OFFSET + 1, rec2, use_grad=False),
                              iso.texture_sample(sc.SCENE_TEXTURE_OFFSET + 1, rec2, use_grad=False))
        # images with the anisotropic sampler on both sides
        p = default_scene.default_params(bounce_count=8)
        r.on_resize(256, 256)
        r.render(8, params=p)
        img = r.read_accumulation()
        ref, _ = ora.render(p, 256, 256, 0, 8)
        assert metrics.close_fraction(img, ref, 1e-4) > 0.99
        assert metrics.rel_mse(img / 8, ref / 8) <= 1e-3
        ref_iso, _ = iso.render(p, 256, 256, 0, 8)
        assert metrics.close_fraction(img, ref_iso, 1e-4) < 0.99  # and it is not the isotropic image
        with pytest.raises(core.PtError):
            r.set_sampler(0)


def test_texture_size_limits(default_scene, oracle_mod):
    """Tuning keys max_texture_size / texture_budget_mb = TextureUploader's limits: the down-scaled textures and their mip
    chains equal the oracle's (blit arithmetic bit for bit), through the sampler probe."""
    rec = _records(np.random.default_rng(5), 2000)
    for kwargs, extent0 in (({"max_texture_size": 512}, (512, 512)), ({"texture_budget_mb": 4}, (256, 256))):
        ora = oracle_mod.OracleScene(default_scene, **kwargs)
        with core.Renderer(0) as r:
            for k, v in kwargs.items():
                r.set_tuning(k, v)
            r.update_scene_data(default_scene)
            assert ora.texture_info(sc.SCENE_TEXTURE_OFFSET)[:2] == extent0
            for slot in range(sc.SCENE_TEXTURE_OFFSET, sc.SCENE_TEXTURE_OFFSET + len(default_scene.textures)):
                _check(r, ora, slot, rec)
