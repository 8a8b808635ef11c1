"""skinning.comp in the oracle: identity bones reproduce the bind pose, a rigid bone is the rigid transform,
blended bones interpolate positions; normals stay unit length."""
import copy
import importlib

import numpy as np

import conftest

sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def _ray_along_x(y, z=0.013):
    r = np.zeros(1, sc.RAY)
    r["origin"], r["direction"] = (-5.0, y, z), (1, 0, 0)
    r["tmin"], r["tmax"] = 1e-5, 1e4
    return r


def test_identity_bones_equal_a_static_copy(oracle_mod):
    s = scenes.skinned_scene()
    static = copy.copy(s)
    # the same tube as a static geometry: animated buffers appended to the static ones by hand
    static.vertices = np.concatenate([s.vertices, np.array([tuple(v[f] for f in sc.VERTEX.names) for v in s.animated_vertices], sc.VERTEX)])
    static.indices = np.concatenate([s.indices, s.animated_indices])
    g = s.geometries.copy()
    g[0]["vertex_offset"], g[0]["index_offset"] = len(s.vertices), len(s.indices)
    static.geometries, static.geometry_is_animated = g, None
    p = s.default_params()
    a = oracle_mod.OracleScene(s).first_hit_aov(p, 96, 72)
    b = oracle_mod.OracleScene(static).first_hit_aov(p, 96, 72)
    # (weights w_i with sum 1 blend P into sum(w_i * P): equal to P up to rounding only)
    same = (a["instance"] == b["instance"]) & (a["primitive"] == b["primitive"])
    assert same.mean() > 0.999
    assert np.allclose(a["t"][same], b["t"][same], rtol=1e-5)


def test_rigid_bone_moves_the_mesh_rigidly(oracle_mod):
    s = scenes.skinned_scene()
    lifted = copy.copy(s)
    lifted.bone_transforms = np.tile(np.array([1, 0, 0, 0.5, 0, 1, 0, 0.25, 0, 0, 1, 0], np.float32), (4, 1))
    # a horizontal ray at height 1 meets the tube's wall; every bone shifts the mesh by (0.5, 0.25, 0): the same
    # point of the wall is met by the ray at height 1.25, half a unit later
    before = oracle_mod.OracleScene(s).trace_closest(_ray_along_x(1.0))
    after = oracle_mod.OracleScene(lifted).trace_closest(_ray_along_x(1.25))
    assert before["instance"][0] == after["instance"][0] == 1 and before["primitive"][0] == after["primitive"][0]
    assert np.isclose(after["t"][0] - before["t"][0], 0.5, atol=1e-5)


def test_bent_pose_changes_hits_only_on_the_skinned_instance(oracle_mod):
    s = scenes.skinned_scene()
    bent = copy.copy(s)
    bent.bone_transforms = scenes.bend_bones(4, 20.0)
    p = s.default_params()
    a = oracle_mod.OracleScene(s).first_hit_aov(p, 96, 72)
    b = oracle_mod.OracleScene(bent).first_hit_aov(p, 96, 72)
    changed = (a["primitive"] != b["primitive"]) | (a["instance"] != b["instance"])
    assert changed.any()
    assert ((a["instance"][changed] == 1) | (b["instance"][changed] == 1)).all()
