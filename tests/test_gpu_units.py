"""CUDA device functions vs the oracle on the reference's unit-test input grids
(Path-Tracing-Tests/ShadingTest.cpp, BsdfTest.cpp, TestData.h) and on random inputs.
Bars: RNG bit-exact; everything else <= 1e-5 relative; plus the reference's own
assertions (finite; lobe weights sum to 1)."""
import numpy as np
import pytest

import refdata as rd

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def noise_floor(oracle_mod, mode, inp, cols=None, trials=6):
    """fp32 conditioning of the ORACLE at these inputs: the largest change of its output when every
    float input moves by at most one ulp.  Where a formula cancels catastrophically (e.g. the
    refraction Jacobian's LdotH + eta * VdotH near 0, or eta ~ 1) two correctly rounded fp32
    implementations cannot agree to 1e-5 of the value; they must agree to within this floor."""
    inp = np.ascontiguousarray(inp, np.float32)
    base = oracle_mod.test_shading(mode, inp).astype(np.float64)
    floor = np.zeros_like(base)
    rs = np.random.default_rng(1000 + mode)
    cols = range(inp.shape[1]) if cols is None else cols
    for _ in range(trials):
        bits = inp.copy().view(np.int32)
        for c in cols:
            live = np.isfinite(inp[:, c]) & (inp[:, c] != 0)
            bits[live, c] += rs.integers(-1, 2, int(live.sum())).astype(np.int32)
        out = oracle_mod.test_shading(mode, bits.view(np.float32)).astype(np.float64)
        d = np.abs(out - base)
        floor = np.maximum(floor, np.where(np.isfinite(d), d, 0))
    return floor


def assert_close(a, b, rtol=RTOL, atol=1e-9, floor=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol + (0 if floor is None else 8 * floor)
    ok = both_nan | (err <= tol)
    assert ok.all(), f"max rel err {np.nanmax(err / (np.abs(b) + atol)):.3e} at {np.argwhere(~ok)[:5].tolist()}"


GRIDS = {
    0: rd.grid_vec3_float,
    1: rd.grid_vec3_float,
    2: rd.grid_vec3_float,
    3: rd.grid_dielectric,
    4: rd.grid_schlick,
    5: rd.grid_reflection,
    6: rd.grid_refraction,
    7: rd.grid_sample_ggx,
    8: rd.grid_lobe_pdfs,
}


@pytest.mark.parametrize("mode", sorted(GRIDS))
def test_reference_grids(renderer, oracle_mod, mode):
    grid = GRIDS[mode]()
    got, want = renderer.test_shading(mode, grid), oracle_mod.test_shading(mode, grid)
    assert np.isfinite(got).all()  # the reference's assertion
    # unit vectors (mode 7) are compared on the scale of the vector, everything else relative
    assert_close(got, want, atol=1e-5 if mode == 7 else 1e-11, floor=noise_floor(oracle_mod, mode, grid))


def test_lobe_pdfs_sum_to_one(renderer):
    out = renderer.test_shading(8, rd.grid_lobe_pdfs())
    assert np.all(np.abs(out.sum(1) - 1.0) <= 4 * np.finfo(np.float32).eps)


def test_refraction_body(renderer, oracle_mod):
    grid = rd.grid_refraction(flip_l=True)
    assert_close(renderer.test_shading(6, grid), oracle_mod.test_shading(6, grid), floor=noise_floor(oracle_mod, 6, grid))


def test_rng_bit_exact(renderer, oracle_mod):
    rs = np.random.default_rng(11)
    inp = rs.integers(0, 4096, (5000, 4)).astype(np.uint32)
    inp[:, 2] = rs.choice([512, 1920, 3840], len(inp))
    inp[:5] = [[1, 0, 512, 0], [0, 1, 512, 0], [255, 255, 512, 0], [0, 0, 512, 1], [1919, 1079, 1920, 7]]
    inp[5] = [0, 0, 1920, 0]
    got = renderer.test_shading(11, inp.view(np.float32)).view(np.uint32)
    want = oracle_mod.test_shading(11, inp.view(np.float32)).view(np.uint32)
    assert (got == want).all()
    assert got[0, 0] == 0x124EA49D and got[0, 1] == 0x1D719993 and (got[5] == 0).all()


def test_random_microfacet(renderer, oracle_mod):
    rs = np.random.default_rng(12)
    n = 4000
    V, L = rd.random_unit_upper(rs, n), rd.random_unit_upper(rs, n)
    a = rs.uniform(0.0001, 1, n).astype(np.float32)
    F = rs.uniform(0, 1, (n, 3)).astype(np.float32)
    eta = rs.choice([1 / 1.5, 1.5, 1.33, 1 / 1.33], n).astype(np.float32)
    Ld = L * np.array([1, 1, -1], np.float32)
    for mode, inp in ((5, np.column_stack([V, L, F, a])), (6, np.column_stack([V, Ld, F, a, eta])),
                      (7, np.column_stack([rs.uniform(0, 1, (n, 2)).astype(np.float32), V, a]))):
        assert_close(renderer.test_shading(mode, inp), oracle_mod.test_shading(mode, inp), atol=1e-5 if mode == 7 else 1e-9,
                     floor=noise_floor(oracle_mod, mode, inp))


def test_evaluate_bsdf(renderer, oracle_mod):
    rs = np.random.default_rng(13)
    n = 6000
    m, V = rd.random_materials(rs, n), rd.random_unit_upper(rs, n)
    L = rd.random_unit_upper(rs, n)
    L[n // 2 :, 2] *= -1  # half of them below the surface: transmission lobe
    inp = np.column_stack([m, V, L])
    assert_close(renderer.test_shading(9, inp), oracle_mod.test_shading(9, inp), atol=1e-9, floor=noise_floor(oracle_mod, 9, inp))


def test_sample_bsdf_and_rng_consumption(renderer, oracle_mod):
    rs = np.random.default_rng(14)
    n = 6000
    m, V = rd.random_materials(rs, n), rd.random_unit_upper(rs, n)
    seeds = rs.integers(1, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    inp = np.column_stack([m, V, seeds.view(np.float32)])
    got, want = renderer.test_shading(10, inp), oracle_mod.test_shading(10, inp)
    # the advanced RNG state must be identical: same lobe, same number of draws
    same_lobe = got[:, 7].view(np.uint32) == want[:, 7].view(np.uint32)
    assert same_lobe.mean() > 0.999  # a draw within 1 ulp of the lobe threshold may flip
    # TIR yields NaN directions on both sides (Q12)
    assert (np.isnan(got[:, 0]) == np.isnan(want[:, 0]))[same_lobe].all()
    ok = same_lobe & ~np.isnan(want[:, 0])
    floor = noise_floor(oracle_mod, 10, inp, cols=range(20))  # not the seed bits
    # unit direction: sqrt(1 - t1^2 - t2^2) in SampleGGX amplifies rounding for samples at the rim of the
    # disk (an error of ~sqrt(eps) that the input-perturbation floor cannot see: u comes from the
    # RNG bits) — 1e-5 for all but a handful, 1e-3 worst case
    derr = np.abs(got[ok, :3].astype(np.float64) - want[ok, :3])
    assert (derr <= 1e-5 + 8 * floor[ok, :3]).mean() > 0.999 and derr.max() <= 1e-3
    # pdf and colour are evaluated AT the sampled direction, so the rim cases above carry over: 1e-5 (+ the
    # conditioning floor) for all but a handful of records, 2e-3 worst case
    a, b = got[ok, 3:7].astype(np.float64), want[ok, 3:7].astype(np.float64)
    err = np.abs(a - b)
    tight = (err <= RTOL * np.abs(b) + 1e-9 + 8 * floor[ok, 3:7]).all(axis=1)
    assert tight.mean() > 0.999, tight.mean()
    assert (err <= 2e-3 * np.abs(b) + 1e-9 + 8 * floor[ok, 3:7]).all()


def test_camera_and_offsets(renderer, oracle_mod, default_scene):
    rs = np.random.default_rng(15)
    n = 2000
    px = rs.integers(0, 1920, n).astype(np.uint32)
    py = rs.integers(0, 1080, n).astype(np.uint32)
    wh = np.tile(np.array([1920, 1080], np.uint32), (n, 1))
    u = rs.uniform(0, 1, (n, 4)).astype(np.float32)
    lens = np.where(np.arange(n) % 2 == 0, 0.0, 0.05).astype(np.float32)
    focal = np.full(n, 4.0, np.float32)
    cam = np.tile(np.concatenate([default_scene.view_inverse, default_scene.proj_inverse]), (n, 1))
    inp = np.column_stack([px.view(np.float32), py.view(np.float32), wh.view(np.float32), u, lens, focal, cam])
    assert_close(renderer.test_shading(12, inp), oracle_mod.test_shading(12, inp), atol=1e-6)
    o = rs.uniform(-10, 10, (n, 3)).astype(np.float32)
    o[: n // 4] *= 1e-3
    nn = rd.random_unit_upper(rs, n) * rs.choice([-1, 1], (n, 1)).astype(np.float32)
    inp = np.column_stack([o, nn])
    assert (renderer.test_shading(13, inp).view(np.uint32) == oracle_mod.test_shading(13, inp).view(np.uint32)).all()
    d = rs.uniform(0, 1, (n, 2)).astype(np.float32)
    assert_close(renderer.test_shading(14, d), oracle_mod.test_shading(14, d), atol=1e-7)
    assert_close(renderer.test_shading(15, nn), oracle_mod.test_shading(15, nn), atol=1e-7)


# ---- the units the reference's tests do not cover: tracing.glsl, ray.glsl:109-131, sampling.glsl, material.glsl ----
# (the ORACLE's functions are bit-identical to the reference's compiled GLSL on the same generators:
# tests/test_oracle_vs_glsl.py)

import unit_inputs as ui


@pytest.mark.parametrize("mode", [16, 17, 18, 19, 20, 21, 22, 23, 24, 25], ids=[ui.MODE_NAMES[m] for m in range(16, 26)])
def test_geometry_and_differential_units(renderer, oracle_mod, mode):
    n = 2000 if mode == 22 else 20000  # one 3 KB light block per sampleLight record
    x = ui.inputs(mode, n, seed=21)
    got, want = renderer.test_shading(mode, x), oracle_mod.test_shading(mode, x)
    assert (np.isnan(got) == np.isnan(want)).all()
    if mode in (16, 17, 21, 24, 25):
        # same operations in the same order: bit for bit
        assert ui.bit_equal(got, want).all(), f"{(~ui.bit_equal(got, want)).sum()} records differ"
        return
    floor = noise_floor(oracle_mod, mode, x)
    if mode == 23:
        # the core transforms with the precomposed (instance x mesh) matrix and normal matrix (k_bake); the GLSL
        # inverts per vertex: positions bit for bit, unit vectors to 1e-5 of their length
        assert ui.bit_equal(got[:, :3], want[:, :3]).all()
        assert_close(got[:, 3:], want[:, 3:], atol=1e-5, floor=floor[:, 3:])
        return
    # unit directions / clamped derivatives: relative 1e-5 on the scale of the value, plus the conditioning floor
    assert_close(got, want, atol=1e-6, floor=floor)
