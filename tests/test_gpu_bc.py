"""Block-compressed textures through the C ABI: GPU decode + sampling against the oracle on the feature scene
with its colour (BC1 sRGB), normal (BC5) and alpha-tested leaf (BC3) textures stored as .dds-style mip chains."""
import copy
import importlib

import numpy as np
import pytest

import bc_ref
import conftest
import metrics

pytestmark = pytest.mark.gpu
core = conftest.core
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def _compressed_feature_scene(levels):
    scene = scenes.feature_scene(texture_size=64)
    s = copy.copy(scene)
    tex = list(scene.textures)
    for i, fmt, srgb in ((0, bc_ref.BC1, True), (1, bc_ref.BC5, False), (3, bc_ref.BC3, True)):
        px = np.ascontiguousarray(tex[i].pixels)
        tex[i] = sc.Texture(bc_ref.encode_chain(fmt, px, levels), srgb=srgb, bc_format=fmt, bc_extent=(px.shape[1], px.shape[0]),
                            levels=levels)
    s.textures = tex
    return s


@pytest.mark.parametrize("levels", [1, 3, 5])
def test_feature_scene_with_bc_textures(oracle_mod, levels):
    s = _compressed_feature_scene(levels)
    p = s.default_params(bounce_count=6)
    W, H = 160, 120
    ora = oracle_mod.OracleScene(s)
    with core.Renderer(0) as r:
        r.update_scene_data(s)
        a, b = r.first_hit_aov(p, W, H), ora.first_hit_aov(p, W, H)
        # the leaf cards are alpha-tested against the BC3 alpha block: same accepted hits
        assert np.array_equal(a["primitive"], b["primitive"]) and np.array_equal(a["instance"], b["instance"])
        r.on_resize(W, H)
        r.render(8, params=p)
        img = r.read_accumulation()
    ref, _ = ora.render(p, W, H, 0, 8)
    assert metrics.close_fraction(img, ref, 1e-4) > 0.99
    assert metrics.rel_mse(img / 8, ref / 8) <= 1e-3


def test_bc_differs_from_uncompressed_only_slightly():
    """Sanity of the test encoder + decoders: the compressed scene renders close to the original one."""
    plain, comp = scenes.feature_scene(texture_size=64), _compressed_feature_scene(5)
    p = plain.default_params(bounce_count=2)
    imgs = []
    for s in (plain, comp):
        with core.Renderer(0) as r:
            r.update_scene_data(s)
            r.on_resize(96, 72)
            r.render(16, params=p)
            imgs.append(r.read_accumulation()[..., :3] / 16)
    assert abs(imgs[0].mean() - imgs[1].mean()) < 0.05 * imgs[0].mean()


def test_rejects_bad_descriptors(default_scene):
    s = copy.copy(default_scene)
    s.textures = list(default_scene.textures) + [sc.Texture(np.zeros(8, np.uint8), bc_format=sc.TEXTURE_BC1, bc_extent=(4, 4), levels=9)]
    with core.Renderer(0) as r:
        with pytest.raises(core.PtError):
            r.update_scene_data(s)
