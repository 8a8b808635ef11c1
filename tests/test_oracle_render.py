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
