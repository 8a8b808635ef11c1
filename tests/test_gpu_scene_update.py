"""pt_scene_update (animated instances / lights: re-bake + BVH rebuild) against a fresh upload of the moved
scene and against the oracle."""
import copy
import importlib

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
core = conftest.core
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def _moved(scene, seed):
    """The scene with every instance rotated about y, scaled and shifted a little (3x4 row-major)."""
    rs = np.random.default_rng(seed)
    moved = copy.copy(scene)
    inst = scene.instances.copy()
    for k in range(len(inst)):
        m = np.eye(4, dtype=np.float64)
        m[:3, :] = inst["transform"][k].reshape(3, 4)
        a = rs.uniform(-0.4, 0.4)
        rot = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]])
        shift = np.eye(4)
        shift[:3, 3] = rs.uniform(-0.15, 0.15, 3)
        inst["transform"][k] = (shift @ m @ rot @ np.diag([1, rs.uniform(0.8, 1.2), 1, 1]))[:3, :].astype(np.float32).reshape(12)
    moved.instances = inst
    return moved


def _same_hits(a, b):
    return all(np.array_equal(a[f], b[f]) for f in ("instance", "geometry", "primitive", "t", "u", "v"))


def test_moved_instances_equal_a_fresh_upload(default_scene, default_oracle, oracle_mod):
    p = default_scene.default_params(bounce_count=6)
    W, H = 160, 120
    moved = _moved(default_scene, 5)
    with core.Renderer(0) as fresh, core.Renderer(0) as upd:
        fresh.update_scene_data(moved)
        upd.update_scene_data(default_scene)
        before = upd.first_hit_aov(p, W, H)
        upd.update_scene(instance_transforms=moved.instances["transform"])
        assert upd.stats()["bvh_build_ms"] > 0
        a, b = fresh.first_hit_aov(p, W, H), upd.first_hit_aov(p, W, H)
        assert _same_hits(a, b) and not _same_hits(before, b)
        ora = oracle_mod.OracleScene(moved).first_hit_aov(p, W, H)
        assert np.array_equal(b["primitive"], ora["primitive"]) and np.array_equal(b["instance"], ora["instance"])
        for r in (fresh, upd):
            r.on_resize(W, H)
            r.render(4, params=p)
        assert np.array_equal(fresh.read_accumulation(), upd.read_accumulation())
        # and back again: the original scene's image, bit for bit
        upd.update_scene(instance_transforms=default_scene.instances["transform"])
        assert _same_hits(before, upd.first_hit_aov(p, W, H))


def test_animation_loop_on_a_tessellated_scene():
    """A few frames of a node animation on a scene large enough for the multi-pass PLOC build."""
    scene = scenes.chess_scene(256, 144, segments=24, rings=20, board_tess=16, texture_size=64)
    p = scene.default_params(bounce_count=4)
    with core.Renderer(0) as upd:
        upd.update_scene_data(scene)
        for frame in range(3):
            moved = _moved(scene, 100 + frame)
            upd.update_scene(instance_transforms=moved.instances["transform"])
            upd.on_resize(256, 144)
            upd.render(2, params=p)
            img = upd.read_accumulation()
            with core.Renderer(0) as fresh:
                fresh.update_scene_data(moved)
                fresh.on_resize(256, 144)
                fresh.render(2, params=p)
                assert np.array_equal(img, fresh.read_accumulation()), frame


def test_light_update(default_scene):
    p = default_scene.default_params(bounce_count=4)
    lit = copy.copy(default_scene)
    dl = np.zeros((), sc.DIRECTIONAL_LIGHT)
    dl["color"], dl["direction"] = (4.0, 3.5, 3.0), (-0.3, -1.0, 0.2)
    pl = np.zeros(2, sc.POINT_LIGHT)
    pl["color"] = [(6, 2, 2), (2, 2, 6)]
    pl["position"] = [(0.5, 1.5, 0.5), (-1.0, 0.8, -0.6)]
    pl["attenuation_constant"], pl["attenuation_quadratic"] = 1.0, 0.2
    lit.directional_light, lit.point_lights = dl, pl
    with core.Renderer(0) as fresh, core.Renderer(0) as upd:
        fresh.update_scene_data(lit)
        upd.update_scene_data(default_scene)
        upd.update_scene(point_lights=pl, directional_light=dl)
        for r in (fresh, upd):
            r.on_resize(96, 96)
            r.render(4, params=p)
        a, b = fresh.read_accumulation(), upd.read_accumulation()
        assert np.array_equal(a, b)
        upd.update_scene(point_lights=default_scene.point_lights, directional_light=default_scene.directional_light)
        upd.on_resize(96, 96)
        upd.render(4, params=p)
        assert not np.array_equal(a, upd.read_accumulation())


def test_errors(default_scene):
    import ctypes as C

    L = core.lib()
    with core.Renderer(0) as r:
        d = core.SceneUpdateDesc()
        assert L.pt_scene_update(r._h, C.addressof(d)) == -5  # PT_ERR_NO_SCENE
        assert L.pt_scene_update(r._h, None) == -1
        r.update_scene_data(default_scene)
        with pytest.raises(core.PtError):
            r.update_scene(instance_transforms=np.zeros((len(default_scene.instances) + 1, 12), np.float32))
        with pytest.raises(core.PtError):
            r.update_scene(point_lights=np.zeros(65, sc.POINT_LIGHT))
        r.update_scene()  # nothing to change is fine
