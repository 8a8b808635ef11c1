"""The procedural stand-ins of BASELINE.json configs[1..4]: declared triangle counts, determinism,
and a finite oracle render of each (CPU only)."""
import numpy as np
import pytest

import conftest  # noqa: F401  (sys.path)
from scenes_small import SMALL, W, H


@pytest.mark.parametrize("name", sorted(SMALL))
def test_declared_counts_and_determinism(name):
    build, _, triangles = SMALL[name]
    a, b = build(), build()
    assert a.instanced_triangle_count() == triangles
    assert (a.vertices.tobytes() == b.vertices.tobytes()) and (a.indices == b.indices).all()
    assert all((x.pixels == y.pixels).all() for x, y in zip(a.textures, b.textures))
    assert a.indices.max() < len(a.vertices)


@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_renders_finite(name, oracle_mod):
    build, bounces, _ = SMALL[name]
    s = build()
    o = oracle_mod.OracleScene(s)
    p = s.default_params(bounce_count=bounces)
    img, cnt = o.render(p, W, H, 0, 2)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.01
    assert cnt["samples"] == W * H * 2 and cnt["hits"] > 0.5 * W * H
    if name == "atrium":  # alpha-tested foliage is actually reached, by both ray types
        assert cnt["alpha_tests_closest"] > 0 and cnt["alpha_tests_shadow"] > 0
    if name == "street":  # 64 point lights + directional
        assert len(s.point_lights) == 64


def test_default_sizes_are_the_declared_ones():
    """The docstrings' arithmetic for the full-size workloads (nothing is built here)."""
    assert 32 * 2 * 192 * 159 + 2 * 256 * 256 + 48 + 2 == 2_084_914
    assert 2 * 4096 * 64 == 524_288
    assert 4 * 12 * 2 * 256 * 511 + 2 * 2 * 1024 ** 2 + 4 * 12 + 12288 * 512 * 2 == 29_335_600
    assert 48 * 2 * 192 ** 2 + 2 * 768 ** 2 + 4096 * (2 + 2 * 32 * 15) == 8_658_944


def test_exr_writer_round_trip(tmp_path):
    """write_exr: a valid uncompressed scanline OpenEXR 2.0 file (magic, version, header attributes, offset table) that
    keeps every bit of the float image."""
    import importlib
    import struct

    import numpy as np

    core = importlib.import_module("path-tracing_b200.core")
    rs = np.random.default_rng(3)
    img = rs.normal(0, 10, (7, 13, 4)).astype(np.float32)
    img[0, 0] = [np.inf, 1e-38, -0.0, 1.0]
    path = tmp_path / "a.exr"
    core.write_exr(str(path), img)
    raw = path.read_bytes()
    assert struct.unpack_from("<ii", raw, 0) == (20000630, 2) and b"channels\0chlist\0" in raw and b"dataWindow\0box2i\0" in raw
    back = core.read_exr(str(path))
    assert back.dtype == np.float32 and np.array_equal(back.view(np.uint32), img.view(np.uint32))
