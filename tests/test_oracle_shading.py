"""The oracle's shading.glsl / bsdf.glsl restatement against (a) the reference tests' own
assertions on the reference tests' own input grids (finite; lobe weights sum to 1) and (b) fp64
closed forms of the same formulas, incl. the sanity values listed in SURVEY §8c."""
import numpy as np
import pytest

import refdata as rd

PI = np.float64(np.float32(3.14159265359))


def finite(a):
    return np.isfinite(a).all()


# ---- fp64 closed forms (numpy) of Path-Tracing/Shaders/shading.glsl -------------------------------
def ggx_d(H, a):
    a2 = a * a
    den = PI * a2 * (H[0] ** 2 / a2 + H[1] ** 2 / a2 + H[2] ** 2) ** 2
    return 1.0 / max(den, 1.0)


def lam(V, a):
    return (np.sqrt(1 + (a * a * V[0] ** 2 + a * a * V[1] ** 2) / V[2] ** 2) - 1) / 2


def smith(V, a):
    return 1 / (1 + lam(V, a))


def fresnel(c, eta):
    s2 = eta * eta * (1 - c * c)
    if s2 > 1:
        return 1.0
    ct = np.sqrt(max(1 - s2, 0))
    rs = (eta * ct - c) / (eta * ct + c)
    rp = (eta * c - ct) / (eta * c + ct)
    return (rs * rs + rp * rp) / 2


def eval_reflection(V, L, F, a):
    if L[2] < 0.00001:
        return np.zeros(3), 0.0
    H = (V + L) / np.linalg.norm(V + L)
    vh = V @ H
    D, Gv, Gl = ggx_d(H, a), smith(V, a), smith(L, a)
    return D * Gv * Gl * F / (4 * V[2]), Gv * max(vh, 0) * D / V[2] / (4 * vh)


def eval_refraction(V, L, F, a, eta):
    if L[2] > -0.00001:
        return np.zeros(3), 0.0
    H = eta * V + L
    H = H / np.linalg.norm(H)
    if H[2] < 0:
        H = -H
    vh, lh = V @ H, L @ H
    D, Gv, Gl = ggx_d(H, a), smith(V, a), smith(L, a)
    jac = eta * eta * abs(lh) / (lh + eta * vh) ** 2
    return abs(vh) / abs(V[2]) * D * Gv * Gl * F * jac, Gv * abs(vh) * D / V[2] * jac


def relclose(a, b, tol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1e-30) + 1e-30)


def test_reference_grids_are_finite(oracle_mod):
    """The assertion the reference makes (Path-Tracing-Tests/TestCommon.h:9-19)."""
    o = oracle_mod.test_shading
    assert finite(o(0, rd.grid_vec3_float())) and finite(o(1, rd.grid_vec3_float())) and finite(o(2, rd.grid_vec3_float()))
    assert finite(o(3, rd.grid_dielectric())) and finite(o(4, rd.grid_schlick()))
    assert finite(o(5, rd.grid_reflection())) and finite(o(6, rd.grid_refraction())) and finite(o(7, rd.grid_sample_ggx()))
    assert rd.grid_reflection().shape == (54, 10) and rd.grid_refraction().shape == (108, 11)
    assert rd.grid_sample_ggx().shape == (24, 6) and rd.grid_lobe_pdfs().shape == (125, 3)


def test_lobe_pdfs_sum_to_one(oracle_mod):
    """BsdfTest.SampleLobePdfs: ASSERT_FLOAT_EQ(D + G + M + T, 1) for the 125 cases (4 ulp)."""
    out = oracle_mod.test_shading(8, rd.grid_lobe_pdfs())
    s = out[:, 0] + out[:, 1] + out[:, 2] + out[:, 3]
    assert np.all(np.abs(s - 1.0) <= 4 * np.finfo(np.float32).eps)


def test_fp64_sanity_values(oracle_mod):
    """fp64 values the fp32 code must reproduce to 1e-5 relative."""
    o = oracle_mod.test_shading
    f = o(3, np.array([[1, 1 / 1.5], [0.5, 1 / 1.5], [0.001, 0.001], [0.999, 0.999], [0.001, 0.999], [0.999, 0.001]], np.float32))[:, 0]
    assert relclose(f[:2], [0.04, 0.0891867128]) and relclose(f[2], 0.499998, 1e-5)
    assert relclose(f[3], 2.5025119e-07, 2e-3)  # catastrophic cancellation at eta ~ 1: fp32 keeps 3 digits
    assert relclose(f[4:], [0.91442693, 0.996007986])
    s = o(4, np.array([[0.001], [0.999]], np.float32))[:, 0]
    assert relclose(s[0], 0.99500999) and relclose(s[1], 1e-15, 1e-3)
    h1, h2 = rd.EDGE_VEC3S[2], rd.EDGE_VEC3S[0]
    d = o(0, np.array([[*h1, 0.001], [*h1, 0.999], [*h2, 0.999], [*h2, 0.001]], np.float32))[:, 0]
    assert d[0] == 1.0 and relclose(d[1:], [0.318947332, 0.317673714, 3.18374844e-07], 2e-5)
    l = o(1, np.array([[*h2, 0.999], [*h1, 0.999]], np.float32))[:, 0]
    assert relclose(l[0], 48.9530277, 2e-5) and relclose(l[1], 2.54559636e-05, 5e-3)
    g = o(2, np.array([[*h2, 0.999]], np.float32))[:, 0]
    assert relclose(g[0], 0.0200188066, 2e-5)


def test_against_fp64_closed_forms(oracle_mod):
    o = oracle_mod.test_shading
    rs = np.random.default_rng(7)
    n = 400
    V, L = rd.random_unit_upper(rs, n), rd.random_unit_upper(rs, n)
    a = rs.uniform(0.02, 1, n).astype(np.float32)
    F = rs.uniform(0, 1, (n, 3)).astype(np.float32)
    eta = rs.choice([1 / 1.5, 1.5, 1.33], n).astype(np.float32)
    refl = o(5, np.column_stack([V, L, F, a]))
    Ld = L * np.array([1, 1, -1], np.float32)
    refr = o(6, np.column_stack([V, Ld, F, a, eta]))
    for i in range(n):
        f, pdf = eval_reflection(V[i].astype(np.float64), L[i].astype(np.float64), F[i].astype(np.float64), np.float64(a[i]))
        assert relclose(refl[i, :3], f, 2e-4) and relclose(refl[i, 3], pdf, 2e-4)
        f, pdf = eval_refraction(V[i].astype(np.float64), Ld[i].astype(np.float64), F[i].astype(np.float64), np.float64(a[i]), np.float64(eta[i]))
        assert relclose(refr[i, :3], f, 5e-4) and relclose(refr[i, 3], pdf, 5e-4)
    d = o(0, np.column_stack([V, a]))[:, 0]
    g = o(2, np.column_stack([V, a]))[:, 0]
    for i in range(n):
        assert relclose(d[i], ggx_d(V[i].astype(np.float64), np.float64(a[i])), 1e-4)
        assert relclose(g[i], smith(V[i].astype(np.float64), np.float64(a[i])), 1e-4)


def test_refraction_body_is_exercised(oracle_mod):
    """The reference grid never reaches EvaluateRefraction's body (L.z > 0); the flipped grid does."""
    out = oracle_mod.test_shading(6, rd.grid_refraction(flip_l=True))
    assert np.isfinite(out).all() and (out[:, 3] > 0).any()
    assert (oracle_mod.test_shading(6, rd.grid_refraction())[:, 3] == 0).all()


def test_sample_ggx_is_unit_upper_hemisphere(oracle_mod):
    h = oracle_mod.test_shading(7, rd.grid_sample_ggx())
    assert np.allclose(np.linalg.norm(h, axis=1), 1, atol=1e-5) and (h[:, 2] >= 0).all()


def test_sample_bsdf_consumes_3_to_7_randoms(oracle_mod):
    """the lobe choice consumes 3, 4, 5 or 7 numbers."""
    rs = np.random.default_rng(3)
    n = 2000
    m = rd.random_materials(rs, n)
    V = rd.random_unit_upper(rs, n)
    seeds = rs.integers(1, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    out = oracle_mod.test_shading(10, np.column_stack([m, V, seeds.view(np.float32)]))
    end = out[:, 7].view(np.uint32)

    def advance(s, k):
        s = s.copy()
        for _ in range(k):
            s ^= (s << np.uint32(13)) & np.uint32(0xFFFFFFFF)
            s ^= s >> np.uint32(17)
            s ^= (s << np.uint32(5)) & np.uint32(0xFFFFFFFF)
        return s

    counts = np.zeros(n, int)
    for k in (3, 4, 5, 7):
        counts[advance(seeds, k) == end] = k
    assert (counts > 0).all() and set(np.unique(counts)) == {3, 4, 5, 7}


@pytest.mark.parametrize("mode,grid", [(14, None), (15, None), (13, None)])
def test_helpers_finite(oracle_mod, mode, grid):
    rs = np.random.default_rng(mode)
    if mode == 14:
        inp = rs.uniform(0, 1, (256, 2)).astype(np.float32)
        out = oracle_mod.test_shading(mode, inp)
        assert (np.linalg.norm(out, axis=1) <= 1 + 1e-6).all()
    elif mode == 15:
        n = rd.random_unit_upper(rs, 256)
        out = oracle_mod.test_shading(mode, n).reshape(-1, 3, 3)
        eye = np.einsum("nij,nkj->nik", out, out)
        assert np.allclose(eye, np.eye(3), atol=1e-5)
    else:
        o = rs.uniform(-10, 10, (256, 3)).astype(np.float32)
        o[:16] *= 1e-3
        n = rd.random_unit_upper(rs, 256)
        out = oracle_mod.test_shading(mode, np.column_stack([o, n]))
        d = out - o
        assert np.isfinite(out).all() and (np.einsum("ij,ij->i", d, n) > 0).all()
