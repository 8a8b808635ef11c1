"""Software traceRayEXT on the GPU (GPU-built BVH4 + watertight triangles + any-hit alpha tests)
against the oracle (CPU SAH BVH2).  Bars: primitive ids bit-exact for rays not grazing an edge,
t within 1e-5 relative."""
import importlib

import numpy as np
import pytest

import conftest
from test_oracle_traversal import random_rays

pytestmark = pytest.mark.gpu
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def compare_hits(a, b, rays=None):
    same = (a["instance"] == b["instance"]) & (a["geometry"] == b["geometry"]) & (a["primitive"] == b["primitive"])
    hit = b["instance"] != sc.NO_HIT
    w = 1 - b["u"] - b["v"]
    grazing = (np.minimum(np.minimum(b["u"], b["v"]), w) < 1e-4) & hit
    wa = 1 - a["u"] - a["v"]
    grazing |= (np.minimum(np.minimum(a["u"], a["v"]), wa) < 1e-4) & (a["instance"] != sc.NO_HIT)
    assert (same | grazing).all(), f"{(~(same | grazing)).sum()} non-grazing primitive mismatches"
    ok = same & hit
    err = np.abs(a["t"][ok].astype(np.float64) - b["t"][ok])
    scale = 1.0 if rays is None else np.abs(rays["origin"][ok.reshape(-1)]).max(axis=1)
    assert (err <= 1e-5 * b["t"][ok] + 4 * np.finfo(np.float32).eps * scale).all()
    # barycentrics: their conditioning is distance / triangle size, so the bound is statistical
    du = np.maximum(np.abs(a["u"][ok] - b["u"][ok]), np.abs(a["v"][ok] - b["v"][ok]))
    if du.size:
        assert np.median(du) <= 1e-5 and du.max() <= 5e-3
    return float(same.mean())


def test_first_hit_aov_default_512(default_renderer, default_oracle, default_scene):
    """BASELINE.json configs[0]: Default scene, 512x512, pixel-centre primary rays."""
    p = default_scene.default_params()
    a, b = default_renderer.first_hit_aov(p, 512, 512), default_oracle.first_hit_aov(p, 512, 512)
    assert compare_hits(a, b) > 0.9999
    assert (b["instance"] != sc.NO_HIT).mean() > 0.85


def test_random_rays_default(default_renderer, default_oracle):
    rays = random_rays(np.random.default_rng(21), 200000, [-7, -1.2, -2.2], [-2, 3.2, 2.2])
    compare_hits(default_renderer.trace_closest(rays), default_oracle.trace_closest(rays), rays)
    rays["tmax"] = np.random.default_rng(22).uniform(0.05, 6, len(rays))
    assert (default_renderer.trace_occlusion(rays) == default_oracle.trace_occlusion(rays)).all()


@pytest.fixture(scope="module")
def feature(oracle_mod):
    s = scenes.feature_scene()
    r = conftest.core.Renderer(0)
    r.update_scene_data(s)
    yield s, r, oracle_mod.OracleScene(s)
    r.close()


def test_feature_scene_alpha_tested(feature):
    """Mesh + instance transforms (incl. non-uniform scale), alpha-tested cards (any-hit)."""
    s, r, o = feature
    p = s.default_params()
    compare_hits(r.first_hit_aov(p, 320, 240), o.first_hit_aov(p, 320, 240))
    rays = random_rays(np.random.default_rng(23), 100000, [-4, 0.2, -4], [4, 3.5, 4])
    compare_hits(r.trace_closest(rays), o.trace_closest(rays), rays)
    rays["tmax"] = np.random.default_rng(24).uniform(0.05, 8, len(rays))
    occ_g, occ_o = r.trace_occlusion(rays), o.trace_occlusion(rays)
    assert (occ_g != occ_o).mean() < 1e-4  # alpha == 1.0 exactly on a texel boundary may flip


def test_bigger_scene_vs_oracle(oracle_mod):
    """~120k triangles: exercises the multi-level LBVH -> BVH4 collapse."""
    s = scenes.chess_scene(640, 360, segments=48, rings=40, board_tess=32, texture_size=64)
    o = oracle_mod.OracleScene(s)
    with conftest.core.Renderer(0) as r:
        r.update_scene_data(s)
        st = r.stats()
        assert st["triangle_count"] == s.instanced_triangle_count() == o.triangle_count
        assert 0 < st["bvh_node_count"] < st["triangle_count"]
        p = s.default_params()
        compare_hits(r.first_hit_aov(p, 640, 360), o.first_hit_aov(p, 640, 360))
        rays = random_rays(np.random.default_rng(25), 100000, [-6, 0.0, -6], [6, 4, 6])
        compare_hits(r.trace_closest(rays), o.trace_closest(rays), rays)


def test_edge_cases(oracle_mod):
    """Empty scene, single triangle, <= 4 triangles (single-leaf root), degenerate triangles."""
    b = scenes.SceneBuilder()
    empty = b.build(scenes.camera_matrices((0, 0, -3), (0, 0, 1), 8, 8), (8, 8))
    with conftest.core.Renderer(0) as r:
        r.update_scene_data(empty)
        assert (r.first_hit_aov(empty.default_params(), 8, 8)["instance"] == sc.NO_HIT).all()
        r.on_resize(8, 8)
        r.render(2, params=empty.default_params())
        img = r.read_accumulation()
        assert np.allclose(img[..., :3], 2 * np.array([0.08, 0.09, 0.1], np.float32)) and (img[..., 3] == 1).all()

        for ntri in (1, 2, 3, 4, 5, 9):
            b = scenes.SceneBuilder()
            rs = np.random.default_rng(ntri)
            v = np.zeros(3 * ntri, sc.VERTEX)
            v["position"] = rs.uniform(-1, 1, (3 * ntri, 3))
            v["normal"] = (0, 0, -1)
            v["tangent"], v["bitangent"] = (1, 0, 0), (0, 1, 0)
            if ntri == 9:
                v["position"][3:6] = v["position"][3]  # a degenerate (point) triangle
            g = b.add_geometry(v, np.arange(3 * ntri, dtype=np.uint32))
            b.add_instance(b.add_model([(g, b.add_material_mr(), None)]))
            s = b.build(scenes.camera_matrices((0, 0, -3), (0, 0, 1), 64, 64), (64, 64))
            r.update_scene_data(s)
            o = oracle_mod.OracleScene(s)
            compare_hits(r.first_hit_aov(s.default_params(), 64, 64), o.first_hit_aov(s.default_params(), 64, 64))


@pytest.mark.parametrize("builder", [0, 1])
def test_pathological_geometry_builds_and_renders(builder):
    """Many identical triangles, zero-area triangles, a NaN vertex and one huge triangle among tiny ones:
    both BVH builders terminate (PLOC's pair order guarantees progress on equal and on non-finite boxes) and
    the finite part of the scene is still hit."""
    rs = np.random.default_rng(4)
    n = 3000
    v = np.zeros(3 * n, sc.VERTEX)
    tri = np.array([[-0.2, -0.2, 0.0], [0.2, -0.2, 0.0], [0.0, 0.2, 0.0]], np.float32)
    v["position"] = np.tile(tri, (n, 1))                       # n copies of one triangle
    v["position"][300:600] = 0.5                                # 100 point triangles
    v["position"][600:900] += rs.uniform(-1, 1, (100, 1, 3)).repeat(3, 1).reshape(300, 3)  # scattered small ones
    v["position"][900:903] = [[-50, -50, 5], [50, -50, 5], [0, 80, 5]]                    # a huge backdrop
    v["position"][903, 0] = np.nan                              # a NaN vertex
    v["normal"] = (0, 0, -1)
    v["tangent"], v["bitangent"] = (1, 0, 0), (0, 1, 0)
    b = scenes.SceneBuilder()
    g = b.add_geometry(v, np.arange(3 * n, dtype=np.uint32))
    b.add_instance(b.add_model([(g, b.add_material_mr(), None)]))
    s = b.build(scenes.camera_matrices((0, 0, -3), (0, 0, 1), 64, 64), (64, 64))
    with conftest.core.Renderer(0) as r:
        r.set_tuning("bvh_builder", builder)
        r.update_scene_data(s)
        aov = r.first_hit_aov(s.default_params(), 64, 64)
        assert (aov["instance"] != sc.NO_HIT).mean() > 0.9     # the backdrop fills the frame
        assert aov["primitive"][32, 32] < n and aov["t"][32, 32] == pytest.approx(3.0, abs=1e-3)  # a copy at z = 0: the lowest id wins
        assert aov["primitive"][32, 32] == 0
        r.on_resize(64, 64)
        r.render(2, params=s.default_params(bounce_count=3))
        assert np.isfinite(r.read_accumulation()).all()
