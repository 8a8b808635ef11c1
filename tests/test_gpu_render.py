"""Full hot path on the GPU vs the oracle, same scenes, cameras and (pixel, sample) seeds."""
import importlib

import numpy as np
import pytest

import conftest
import metrics

pytestmark = pytest.mark.gpu
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")


def test_default_scene_config0(default_renderer, default_oracle, default_scene):
    """BASELINE.json configs[0]: Default scene 512x512, 16 spp (TotalSamples 0..15), depth 8."""
    p = default_scene.default_params(bounce_count=8)
    r = default_renderer
    r.on_resize(512, 512)
    r.render(16, params=p)
    img = r.read_accumulation()
    st = r.stats()
    ref, cnt = default_oracle.render(p, 512, 512, 0, 16)
    assert np.isfinite(img).all() and (img[..., 3] == 1).all()
    # per-pixel: a path is a chaotic function of its inputs, so a 1-ulp difference can send a
    # few paths elsewhere; the overwhelming majority of pixels must agree to 1e-4
    assert metrics.close_fraction(img, ref, 1e-4) > 0.995
    # image-level bars of BASELINE.md §5 at matched spp
    assert metrics.rel_mse(img / 16, ref / 16) <= 1e-3
    assert metrics.flip(img / 16, ref / 16) <= 5e-3
    assert abs(img[..., :3].mean() - ref[..., :3].mean()) <= 1e-3 * ref[..., :3].mean()
    # same work: ray and sample counts agree to a handful of diverged paths
    assert st["samples"] == cnt["samples"] == 512 * 512 * 16
    assert abs(st["rays_closest"] - cnt["rays_closest"]) <= 1e-4 * cnt["rays_closest"]
    assert abs(st["hits"] - cnt["hits"]) <= 1e-4 * cnt["hits"]


def test_large_frame_grid_stride_loops(default_renderer, default_oracle, default_scene):
    """A frame with more hits per wavefront iteration than resident threads: the grid-stride loop of
    k_shade (with its software-pipelined record prefetch) and every persistent warp of the traversal
    kernels run several times.  (Smaller frames fit one pass and would not notice a broken loop.)"""
    p = default_scene.default_params(bounce_count=8)
    r = default_renderer
    W, H, spp = 1536, 1024, 2
    r.on_resize(W, H)
    r.render(spp, params=p)
    img = r.read_accumulation()
    st = r.stats()
    ref, cnt = default_oracle.render(p, W, H, 0, spp)
    assert np.isfinite(img).all()
    assert metrics.close_fraction(img, ref, 1e-4) > 0.995
    assert st["samples"] == cnt["samples"] == W * H * spp
    assert abs(st["rays_closest"] - cnt["rays_closest"]) <= 1e-4 * cnt["rays_closest"]
    assert abs(st["hits"] - cnt["hits"]) <= 1e-4 * cnt["hits"]


def test_incremental_equals_batch_and_is_deterministic(default_renderer, default_scene):
    """16 frames of 1 sample == one call of 16 samples; reruns are bit-identical."""
    p = default_scene.default_params()
    r = default_renderer
    r.on_resize(256, 256)
    r.render(6, params=p)
    a = r.read_accumulation().copy()
    r.on_resize(256, 256)
    for _ in range(6):
        r.render(1, params=p)
    b = r.read_accumulation().copy()
    assert r.total_samples == 6
    assert (a == b).all()
    r.on_resize(256, 256)
    r.render(4, params=p)
    r.render(2, params=p)
    assert (r.read_accumulation() == a).all()


def test_tiles_are_bit_identical_to_full_frame(default_renderer, default_scene):
    """Image-tile partitioning (multi-GPU) uses global pixel coordinates for the RNG."""
    p = default_scene.default_params()
    r = default_renderer
    r.on_resize(200, 120)
    r.render(3, params=p)
    full = r.read_accumulation().copy()
    r.on_resize(200, 120)
    tiles = np.array([(0, 0, 200, 37), (0, 37, 99, 120)], sc.TILE)
    r.render(3, params=p, tiles=tiles, first_sample=0)
    part = r.read_accumulation().copy()
    inside = np.zeros((120, 200), bool)
    inside[:37] = True
    inside[37:, :99] = True
    assert (part[inside] == full[inside]).all()
    assert (part[~inside] == 0).all()
    r.render(3, params=p, tiles=np.array([(99, 37, 200, 120)], sc.TILE), first_sample=0)
    assert (r.read_accumulation() == full).all()


def test_sample_slices_sum_to_full(default_renderer, default_scene):
    """Sample-sliced partitioning: [0,3) + [3,8) == [0,8) up to fp32 summation order."""
    p = default_scene.default_params()
    r = default_renderer
    r.on_resize(128, 128)
    r.render(8, params=p)
    full = r.read_accumulation().copy()
    r.on_resize(128, 128)
    r.render(3, params=p, first_sample=0)
    a = r.read_accumulation().copy()
    r.on_resize(128, 128)
    r.render(5, params=p, first_sample=3)
    b = r.read_accumulation().copy()
    assert np.allclose(a[..., :3] + b[..., :3], full[..., :3], rtol=1e-6, atol=1e-6)


@pytest.fixture(scope="module")
def feature(oracle_mod):
    s = scenes.feature_scene()
    r = conftest.core.Renderer(0)
    r.update_scene_data(s)
    yield s, r, oracle_mod.OracleScene(s)
    r.close()


@pytest.mark.parametrize("lens", [0.0, 0.08])
def test_feature_scene(feature, lens):
    """Every material model, textures + mips, alpha-tested decals, transmission with volume
    attenuation, point + directional NEE with shadow rays, thin lens."""
    s, r, o = feature
    p = s.default_params(bounce_count=6)
    p.lens_radius, p.focal_distance = lens, 5.0
    W, H, spp = 160, 120, 32
    r.on_resize(W, H)
    r.render(spp, params=p)
    img = r.read_accumulation()
    st = r.stats()
    ref, cnt = o.render(p, W, H, 0, spp)
    assert np.isfinite(img).all()
    assert st["rays_shadow"] > 0 and cnt["alpha_tests_closest"] > 0
    assert metrics.close_fraction(img, ref, 1e-3) > 0.97
    assert metrics.rel_mse(img / spp, ref / spp) <= 1e-3
    assert metrics.flip(img / spp, ref / spp) <= 5e-3
    assert abs(st["rays_closest"] - cnt["rays_closest"]) <= 2e-3 * cnt["rays_closest"]
    # the core skips occlusion queries whose contribution is exactly zero, the oracle traces them
    assert st["rays_shadow"] <= cnt["rays_shadow"]


def test_traversal_stats_toggle(feature):
    s, r, _ = feature
    p = s.default_params(bounce_count=4)
    r.on_resize(64, 48)
    r.set_traversal_stats(True)
    r.render(2, params=p)
    a, st = r.read_accumulation().copy(), r.stats()
    r.set_traversal_stats(False)
    r.on_resize(64, 48)
    r.render(2, params=p)
    assert (r.read_accumulation() == a).all()  # counting does not change the image
    assert st["box_tests_closest"] > st["tri_tests_closest"] > 0 and st["alpha_tests_closest"] > 0
    assert st["box_tests_shadow"] > 0 and st["texel_fetches"] >= 5 * st["hits"]
    assert r.stats()["box_tests_closest"] == 0


def test_scheduling_knobs_do_not_change_the_image(oracle_mod):
    """Pools, slot count, hit sorting and the per-round sample budget only change scheduling, the BVH builder
    (LBVH / PLOC at any radius) only the tree: the accumulated image is bit-identical (closest hits are
    BVH-independent — ties in t go to the smaller triangle id — and samples are summed in sample order)."""
    s = scenes.chess_scene(320, 180, segments=24, rings=20, board_tess=32, texture_size=128)
    p = s.default_params(bounce_count=6)
    W, H, spp = 320, 180, 12
    images = []
    for knobs in ({}, {"pools": 1, "sort_hits": 0}, {"pools": 3, "slots": 200_000}, {"pools": 8, "slots": 140_000, "sbuf_mb": 4},
                  {"pools": 2, "slots": 5000, "sbuf_mb": 1}, {"bvh_builder": 0}, {"bvh_builder": 1, "ploc_radius": 3},
                  {"bvh_builder": 1, "ploc_radius": 40, "pools": 1}):
        r = conftest.core.Renderer(0)
        try:
            for k, v in knobs.items():
                r.set_tuning(k, v)
            r.update_scene_data(s)
            r.on_resize(W, H)
            r.render(spp, params=p)
            images.append(r.read_accumulation().copy())
            st = r.stats()
            assert st["samples"] >= W * H * spp  # restarts count as samples too
        finally:
            r.close()
    for img in images[1:]:
        assert (img == images[0]).all()
    # and the image is the oracle's
    ref, _ = oracle_mod.OracleScene(s).render(p, W, H, 0, spp)
    assert metrics.close_fraction(images[0], ref, 1e-3) > 0.97
    assert metrics.rel_mse(images[0] / spp, ref / spp) <= 1e-3


def test_kernel_timing(feature):
    s, r, _ = feature
    r.on_resize(64, 48)
    r.set_kernel_timing(True)
    r.render(2, params=s.default_params(bounce_count=4))
    st = r.stats()
    r.set_kernel_timing(False)
    assert all(ms > 0 for ms in st["kernel_ms"].values())
    assert sum(st["kernel_launch_count"].values()) == st["kernel_launches"] - 1  # + k_init
    assert sum(st["kernel_ms"].values()) <= st["last_render_ms"] * 1.05


def test_skybox_2d(oracle_mod):
    b = scenes.SceneBuilder()
    rs = np.random.default_rng(31)
    sky = (rs.uniform(0, 4, (32, 64, 4))).astype(np.float32)
    b.set_skybox_2d(sky)
    g = b.add_geometry(*scenes.sphere(1.0, 24, 12))
    b.add_instance(b.add_model([(g, b.add_material_mr(color=(0.9, 0.9, 0.9, 1), roughness=0.1, metalness=1.0), None)]))
    b.set_directional_light((0, 0, 0), (0, -1, 0))
    s = b.build(scenes.camera_matrices((0, 0.5, -4), (0, -0.1, 1), 96, 64, fov_deg=60), (96, 64))
    o = oracle_mod.OracleScene(s)
    with conftest.core.Renderer(0) as r:
        r.update_scene_data(s)
        r.on_resize(96, 64)
        p = s.default_params(4)
        assert p.miss_flags == sc.MISS_FLAGS_SKYBOX_2D
        r.render(8, params=p)
        img = r.read_accumulation()
        ref, _ = o.render(p, 96, 64, 0, 8)
        assert metrics.close_fraction(img, ref, 1e-3) > 0.98
        assert metrics.rel_mse(img / 8, ref / 8) <= 1e-3


def _cube_scene(rs, srgb):
    b = scenes.SceneBuilder()
    if srgb:
        faces = [rs.integers(0, 256, (16, 16, 4), dtype=np.uint8) for _ in range(6)]
    else:
        faces = [rs.uniform(0, 3, (16, 16, 4)).astype(np.float32) for _ in range(6)]
    b.set_skybox_cube(faces, srgb=srgb)
    g = b.add_geometry(*scenes.sphere(1.0, 24, 12))
    b.add_instance(b.add_model([(g, b.add_material_mr(color=(0.9, 0.9, 0.9, 1), roughness=0.1, metalness=1.0), None)]))
    b.set_directional_light((0, 0, 0), (0, -1, 0))
    return b.build(scenes.camera_matrices((0.3, 0.5, -4), (0, -0.1, 1), 96, 64, fov_deg=90), (96, 64))


@pytest.mark.parametrize("srgb", [False, True])
def test_skybox_cube(oracle_mod, srgb):
    """miss.rmiss:29-32: cube sky, float and sRGB8 faces; the mirror ball reaches all six faces."""
    s = _cube_scene(np.random.default_rng(37), srgb)
    o = oracle_mod.OracleScene(s)
    with conftest.core.Renderer(0) as r:
        r.update_scene_data(s)
        r.on_resize(96, 64)
        p = s.default_params(4)
        assert p.miss_flags == sc.MISS_FLAGS_SKYBOX_CUBE
        r.render(8, params=p)
        img = r.read_accumulation()
        ref, _ = o.render(p, 96, 64, 0, 8)
        assert metrics.close_fraction(img, ref, 1e-3) > 0.98
        assert metrics.rel_mse(img / 8, ref / 8) <= 1e-3
        # without the flag the constant sky is used (Renderer.cpp:679-690 sets the flag from the scene)
        p.miss_flags = sc.MISS_FLAGS_NONE
        r.on_resize(96, 64)
        r.render(1, params=p)
        assert np.allclose(r.read_accumulation()[0, 0, :3], (0.08, 0.09, 0.1))


def test_errors(default_scene):
    core = conftest.core
    with core.Renderer(0) as r:
        with pytest.raises(core.PtError) as e:
            r.on_resize(4, 4)
            r.render(1, params=default_scene.default_params())
        assert e.value.status == -5  # PT_ERR_NO_SCENE
        r.update_scene_data(default_scene)
        bad = default_scene.default_params(bounce_count=0)
        with pytest.raises(core.PtError):
            r.render(1, params=bad)
    with pytest.raises(core.PtError):
        core.Renderer(99)


@pytest.mark.parametrize("samples_per_frame", [2, 4])
def test_samples_per_frame(feature, default_renderer, default_oracle, default_scene, samples_per_frame):
    """pt_render_frames: the reference's Release profile renders SamplesPerFrame > 1 samples per vkCmdTraceRaysKHR —
    one rng stream seeded from TotalSamples and one radiance sum per pixel and frame (raygen.rgen:36-118,
    Renderer.cpp:1688-1700).  The oracle's frame loop is bit-identical to the reference's compiled raygen
    (tests/test_oracle_vs_glsl.py)."""
    for (s, r, o, W, H, bounces) in ((default_scene, default_renderer, default_oracle, 256, 256, 8),) + ((feature[0], feature[1], feature[2], 160, 120, 6),):
        p = s.default_params(bounce_count=bounces)
        frames = 6
        r.on_resize(W, H)
        r.render_frames(frames, samples_per_frame, params=p)
        img = r.read_accumulation()
        st = r.stats()
        ref, cnt = o.render_frames(p, W, H, 0, frames, samples_per_frame)
        n = frames * samples_per_frame
        assert np.isfinite(img).all() and (img[..., 3] == 1).all()
        assert st["samples"] == W * H * n
        assert metrics.close_fraction(img, ref, 1e-3) > 0.97
        assert metrics.rel_mse(img / n, ref / n) <= 1e-3
        assert abs(st["rays_closest"] - cnt["rays_closest"]) <= 2e-3 * cnt["rays_closest"]
        # a frame of S samples is NOT S frames of one sample: different seeds (TotalSamples advances by S per frame)
        r.on_resize(W, H)
        r.render(n, params=p)
        assert not np.array_equal(r.read_accumulation(), img)
        # but with S = 1 the two entry points are the same thing, bit for bit
        single = r.read_accumulation()
        r.on_resize(W, H)
        r.render_frames(n, 1, params=p)
        assert np.array_equal(r.read_accumulation(), single)


def test_samples_per_frame_continues_and_partitions(default_renderer, default_scene):
    """Frames are independent given TotalSamples: two calls == one call; tiles == full frame (bit for bit)."""
    p = default_scene.default_params()
    r = default_renderer
    W = H = 128
    r.on_resize(W, H)
    r.render_frames(4, 3, params=p)
    whole = r.read_accumulation()
    r.on_resize(W, H)
    r.render_frames(2, 3, params=p)
    r.render_frames(2, 3, params=p)  # first_sample continues at 6
    assert np.array_equal(r.read_accumulation(), whole)
    r.on_resize(W, H)
    tiles = np.array([(0, 0, W, 40), (0, 40, 64, H)], sc.TILE)
    r.render_frames(4, 3, params=p, tiles=tiles, first_sample=0)
    part = r.read_accumulation()
    mask = np.zeros((H, W), bool)
    mask[:40] = True
    mask[40:, :64] = True
    assert np.array_equal(part[mask], whole[mask]) and (part[~mask] == 0).all()
    with pytest.raises(conftest.core.PtError):
        r.render_frames(1, 0, params=p)


def test_robustness_of_the_boundary(default_scene):
    """Malformed input fails loudly instead of faulting the device: an index beyond its geometry's vertices,
    overlapping tiles, NULL arrays with non-zero counts; the BVH depth is reported and no traversal overflowed."""
    import copy

    core = conftest.core
    with core.Renderer(0) as r:
        bad = copy.copy(default_scene)
        bad.indices = default_scene.indices.copy()
        bad.indices[5] = 1_000_000
        with pytest.raises(core.PtError) as e:
            r.update_scene_data(bad)
        assert e.value.status == -1 and "vertex_length" in str(e.value)
        r.update_scene_data(default_scene)  # the context is still usable
        p = default_scene.default_params()
        r.on_resize(64, 64)
        with pytest.raises(core.PtError) as e:
            r.render(1, params=p, tiles=np.array([(0, 0, 40, 40), (30, 30, 64, 64)], sc.TILE))
        assert "overlap" in str(e.value)
        with pytest.raises(core.PtError):
            r.render(1, params=p, tiles=np.array([(0, 0, 64, 64)] * 3, sc.TILE))
        r.render(2, params=p, tiles=np.array([(0, 0, 32, 64), (32, 0, 64, 64)], sc.TILE))
        st = r.stats()
        assert 1 <= st["bvh_max_depth"] <= 8 and st["stack_overflows"] == 0
        assert np.isfinite(r.read_accumulation()).all()
