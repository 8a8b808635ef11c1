"""Scenes imported by the reference's OWN SceneImporter (assimp, overlay build) from the glTF
files inside vendor/assimp/test/models/glTF2, committed as fixtures by tests/golden/make_imported_scenes.py."""
import os

import numpy as np
import pytest

import conftest
import metrics

pkg, core = conftest.pkg, conftest.core
GOLDEN = os.path.join(conftest.ROOT, "tests", "golden")
CASES = {"box_textured": (256, 256, 12, 1), "cylinder_engine": (320, 240, 121496, 0)}


@pytest.fixture(scope="module", params=sorted(CASES))
def imported(request):
    return request.param, pkg.SceneData.load_npz(os.path.join(GOLDEN, request.param + ".npz"))


def test_fixture_contents(imported, oracle_mod):
    name, s = imported
    w, h, tris, textures = CASES[name]
    assert s.instanced_triangle_count() == tris and len(s.textures) == textures and s.camera_extent == (w, h)
    assert len(s.mr_materials) >= 2 and s.geometry_is_animated is None
    # the importer's material ids address existing materials, its texture indices existing slots
    assert (s.mesh_records["material_id"] >> 8).max() < len(s.mr_materials)
    assert s.mr_materials["color_idx"].max() < pkg.scene.SCENE_TEXTURE_OFFSET + len(s.textures)
    o = oracle_mod.OracleScene(s)
    assert o.triangle_count == tris
    aov = o.first_hit_aov(s.default_params(), w // 4, h // 4)
    hit = aov["instance"] != 0xFFFFFFFF
    assert 0.05 < hit.mean() < 0.95  # the model is in view, with background around it


@pytest.mark.gpu
def test_gpu_matches_oracle(imported, oracle_mod):
    name, s = imported
    w, h, _, _ = CASES[name]
    p = s.default_params(bounce_count=6)
    ora = oracle_mod.OracleScene(s)
    with core.Renderer(0) as r:
        r.update_scene_data(s)
        a, b = r.first_hit_aov(p, w, h), ora.first_hit_aov(p, w, h)
        # ids are bit-exact except for rays that graze an edge (BASELINE north_star): the only pixel that ever
        # differs is a pixel centre exactly ON the edge between two faces of the axis-aligned box (barycentric
        # v = +-0), where the two triangles' t differ in the last bit
        diff = (a["instance"] != b["instance"]) | (a["geometry"] != b["geometry"]) | (a["primitive"] != b["primitive"])
        assert diff.mean() <= 1e-4, diff.sum()
        on_edge = np.minimum(np.minimum(np.abs(a["u"]), np.abs(a["v"])), np.abs(1 - a["u"] - a["v"])) <= 1e-6
        assert on_edge[diff].all()
        hit = a["instance"] != 0xFFFFFFFF
        assert np.allclose(a["t"][hit], b["t"][hit], rtol=1e-5)
        r.on_resize(w, h)
        r.render(4, params=p)
        img = r.read_accumulation()
    ref, _ = ora.render(p, w, h, 0, 4)
    assert metrics.close_fraction(img, ref, 1e-4) > 0.99, name
    assert metrics.rel_mse(img / 4, ref / 4) <= 1e-3
    assert metrics.flip(img / 4, ref / 4) <= 5e-3
