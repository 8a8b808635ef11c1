"""Skeletal animation through the C ABI: k_skin + re-bake + rebuild against the oracle's restatement of skinning.comp."""
import copy
import importlib

import numpy as np
import pytest

import conftest
import metrics

pytestmark = pytest.mark.gpu
core = conftest.core
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")

W, H = 160, 120


def _posed(scene, angle):
    s = copy.copy(scene)
    s.bone_transforms = scenes.bend_bones(len(scene.bone_transforms), angle)
    return s


def _hits_agree(a, b):
    """Skinned positions carry the last-bit differences of the normal matrices nowhere (positions use the bone
    matrix itself), so ids agree except where a pixel centre grazes an edge; t to 1e-5."""
    same = (a["primitive"] == b["primitive"]) & (a["instance"] == b["instance"])
    assert same.mean() > 0.999, same.mean()
    hit = same & (a["instance"] != 0xFFFFFFFF)
    assert np.allclose(a["t"][hit], b["t"][hit], rtol=1e-5)


@pytest.mark.parametrize("angle", [0.0, 12.0, -25.0])
def test_upload_with_pose_matches_oracle(oracle_mod, angle):
    s = _posed(scenes.skinned_scene(W, H), angle)
    p = s.default_params(bounce_count=6)
    ora = oracle_mod.OracleScene(s)
    with core.Renderer(0) as r:
        r.update_scene_data(s)
        assert r.stats()["triangle_count"] == s.instanced_triangle_count()
        _hits_agree(r.first_hit_aov(p, W, H), ora.first_hit_aov(p, W, H))
        r.on_resize(W, H)
        r.render(8, params=p)
        img = r.read_accumulation()
    ref, _ = ora.render(p, W, H, 0, 8)
    assert metrics.close_fraction(img, ref, 1e-3) > 0.98
    assert metrics.rel_mse(img / 8, ref / 8) <= 1e-3


def test_bone_update_equals_fresh_upload():
    base = scenes.skinned_scene(W, H)
    p = base.default_params(bounce_count=4)
    with core.Renderer(0) as upd:
        upd.update_scene_data(base)
        rest = upd.first_hit_aov(p, W, H)
        for angle in (8.0, 20.0, -15.0):
            posed = _posed(base, angle)
            upd.update_scene(bone_transforms=posed.bone_transforms)
            upd.on_resize(W, H)
            upd.render(2, params=p)
            with core.Renderer(0) as fresh:
                fresh.update_scene_data(posed)
                a, b = fresh.first_hit_aov(p, W, H), upd.first_hit_aov(p, W, H)
                assert all(np.array_equal(a[f], b[f]) for f in ("instance", "primitive", "t", "u", "v"))
                assert not np.array_equal(rest["t"], b["t"])
                fresh.on_resize(W, H)
                fresh.render(2, params=p)
                assert np.array_equal(fresh.read_accumulation(), upd.read_accumulation())
        # bones and instances in one update
        moved = base.instances["transform"].copy()
        moved[1, 3] += 0.25
        upd.update_scene(instance_transforms=moved, bone_transforms=scenes.bend_bones(4, 5.0))
        assert not np.array_equal(rest["t"], upd.first_hit_aov(p, W, H)["t"])


def test_errors():
    base = scenes.skinned_scene(W, H)
    with core.Renderer(0) as r:
        r.update_scene_data(base)
        with pytest.raises(core.PtError):
            r.update_scene(bone_transforms=np.zeros((3, 12), np.float32))
    plain = scenes.feature_scene()
    with core.Renderer(0) as r:
        r.update_scene_data(plain)
        with pytest.raises(core.PtError):
            r.update_scene(bone_transforms=np.zeros((4, 12), np.float32))
    bad = copy.copy(base)
    bad.bone_transforms = np.zeros((0, 12), np.float32)
    with core.Renderer(0) as r:
        with pytest.raises(core.PtError):
            r.update_scene_data(bad)
