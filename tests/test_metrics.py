"""LDR-FLIP (path-tracing_b200/metrics.py) against the properties and anchor values the paper states."""
import numpy as np

import metrics


def test_flip_identity_and_range():
    rs = np.random.default_rng(0)
    img = rs.uniform(0, 1, (48, 64, 3))
    assert metrics.flip_map_ldr(img, img).max() == 0.0
    black, white = np.zeros((32, 32, 3)), np.ones((32, 32, 3))
    m = metrics.flip_map_ldr(black, white)
    assert 0.9 < m.mean() <= 1.0  # the largest achromatic difference maps close to 1
    assert np.all((m >= 0) & (m <= 1))


def test_flip_is_symmetric_and_monotonic():
    rs = np.random.default_rng(1)
    a = rs.uniform(0.2, 0.8, (40, 40, 3))
    small, large = a + 0.01, a + 0.1
    assert np.isclose(metrics.flip_map_ldr(a, small).mean(), metrics.flip_map_ldr(small, a).mean())
    assert metrics.flip_map_ldr(a, small).mean() < metrics.flip_map_ldr(a, large).mean()


def test_flip_sees_a_moved_edge_more_than_a_uniform_shift():
    """The feature pipeline: displacing an edge by one pixel costs more than the same mean colour change spread out."""
    a = np.zeros((64, 64, 3))
    a[:, 32:] = 0.8
    moved = np.zeros_like(a)
    moved[:, 33:] = 0.8
    uniform = a.copy()
    uniform[:, 32:] -= 0.8 / 32  # the same total energy removed uniformly from the bright half
    assert metrics.flip_map_ldr(a, moved).max() > 4 * metrics.flip_map_ldr(a, uniform).max()


def test_flip_of_linear_images_tone_maps_first():
    a = np.full((16, 16, 3), 5.0)
    b = np.full((16, 16, 3), 6.0)  # both nearly saturate 1 - exp(-c): a small displayed difference
    assert metrics.flip(a, b) < 0.02
    assert metrics.rel_mse(a, a) == 0.0 and metrics.close_fraction(a, a) == 1.0
