"""Frame / seed schedule semantics of the oracle's raygen restatement (small sizes)."""
import numpy as np

import conftest

sc = conftest.pkg.scene


def test_frames_accumulate_like_the_reference(default_oracle, default_scene):
    """accum += radiance per frame, alpha = 1 (raygen.rgen:115-117); TotalSamples seeds the frame."""
    p = default_scene.default_params(4)
    a, _ = default_oracle.render(p, 48, 48, 0, 3)
    b = np.zeros((48, 48, 4), np.float32)
    for f in range(3):
        default_oracle.render(p, 48, 48, f, 1, accum=b)
    assert (a == b).all() and (a[..., 3] == 1).all()
    c, _ = default_oracle.render(p, 48, 48, 1, 1)
    d, _ = default_oracle.render(p, 48, 48, 0, 1)
    assert not (c == d).all()


def test_tiles_and_threads_do_not_change_pixels(default_oracle, default_scene):
    p = default_scene.default_params(4)
    full, cnt = default_oracle.render(p, 64, 40, 0, 2, threads=1)
    multi, cnt2 = default_oracle.render(p, 64, 40, 0, 2, threads=4)
    assert (full == multi).all() and cnt == cnt2
    tiles = np.array([(0, 0, 64, 13), (10, 13, 50, 40)], sc.TILE)
    part, _ = default_oracle.render(p, 64, 40, 0, 2, tiles=tiles)
    inside = np.zeros((40, 64), bool)
    inside[:13] = True
    inside[13:, 10:50] = True
    assert (part[inside] == full[inside]).all() and (part[~inside] == 0).all()


def test_bounce_limit_and_counters(default_oracle, default_scene):
    p1 = default_scene.default_params(1)
    img, cnt = default_oracle.render(p1, 32, 32, 0, 1)
    assert cnt["rays_closest"] == 32 * 32 and cnt["samples"] == 32 * 32
    assert cnt["rays_shadow"] == cnt["hits"]  # Q11: a shadow ray for every non-miss bounce
    p8 = default_scene.default_params(8)
    _, cnt8 = default_oracle.render(p8, 32, 32, 0, 1)
    assert cnt8["rays_closest"] > cnt["rays_closest"]
    # pixel (0,0) of frame 0 has the all-zero RNG stream (Q2) and still renders
    assert np.isfinite(img[0, 0]).all()


def test_cube_sky_face_selection(oracle_mod):
    """miss.rmiss:29-32 with constant-colour faces: a ray along each axis returns that face's colour
    (Vulkan layer order +X, -X, +Y, -Y, +Z, -Z = Front, Back, Up, Down, Left, Right,
    TextureUploader.cpp:232-236), and the face's (s, t) orientation follows the Vulkan table."""
    import importlib

    scenes = importlib.import_module("path-tracing_b200.scenes")
    colors = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0, 1, 1], [1, 0, 1]], np.float32)
    faces = []
    for c in colors:
        f = np.ones((8, 8, 4), np.float32)
        f[..., :3] = c
        f[0, 0, :3] = 5.0  # texel (s, t) = (0, 0)
        faces.append(f)
    for axis, sign, face in ((0, 1, 0), (0, -1, 1), (1, 1, 2), (1, -1, 3), (2, 1, 4), (2, -1, 5)):
        b = scenes.SceneBuilder()
        b.set_skybox_cube(faces)
        d = np.zeros(3)
        d[axis] = sign
        up = (0, 0, 1) if axis == 1 else (0, 1, 0)
        s = b.build(scenes.camera_matrices((0, 0, 0), d, 9, 9, fov_deg=90, up=up), (9, 9))
        o = oracle_mod.OracleScene(s)
        img, cnt = o.render(s.default_params(2), 9, 9, 0, 4)
        assert cnt["hits"] == 0
        assert np.allclose(img[4, 4, :3] / 4, colors[face], atol=1e-6), (axis, sign)
        # exactly one corner region of the image sees the marked texel
        bright = (img[..., :3].max(-1) / 4 > 1.5)
        assert 0 < bright.sum() <= 16
    # +X face: sc = -z, tc = -y -> texel (0, 0) lies towards +z, +y
    b = scenes.SceneBuilder()
    b.set_skybox_cube(faces)
    s = b.build(scenes.camera_matrices((0, 0, 0), (1, 0.93, 0.93), 3, 3, fov_deg=5), (3, 3))
    img, _ = oracle_mod.OracleScene(s).render(s.default_params(2), 3, 3, 0, 1)
    assert img[1, 1, 0] > 1.5
