"""Seeded input records for every mode of pt_test_shading / pto_test_shading / glr_test_shading
(include/pt_core.h PT_TEST_*), shared by the oracle-vs-compiled-GLSL tests, the golden-vector generator
and the GPU unit tests."""
import numpy as np

import refdata as rd

F = np.float32

MODE_NAMES = [
    "GGXDistribution", "Lambda", "GGXSmith", "DielectricFresnel", "SchlickFresnel", "EvaluateReflection", "EvaluateRefraction",
    "SampleGGX", "sampleLobePdfs", "evaluateBSDF", "sampleBSDF", "initRng_rand", "constructPrimaryRay",
    "offsetRayOriginSelfIntersection", "sampleUniformDiskConcentric", "computeTangentSpace", "computeDpnDuv", "computeDpDxy",
    "computeDerivatives", "computeReflectedDifferentialRays", "computeRefractedDifferentialRays",
    "offsetRayOriginShadowTerminator", "sampleLight", "transform_Vertex", "ReconstructNormalFromXY", "hdrToLdr",
]


def _unit(rs, n):
    v = rs.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(F)


def _vec(rs, n, s=1.0):
    return rs.normal(0, s, (n, 3)).astype(F)


def _bits(a):
    return np.ascontiguousarray(a, np.uint32).view(F)


def _rows34(rs, n, s=0.5):
    """3x4 row-major affine matrices: identity + noise (always well conditioned enough to invert)."""
    m = np.tile(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], F), (n, 1)) + rs.normal(0, s, (n, 12)).astype(F)
    return m.astype(F)


def inputs(mode: int, n: int, seed: int = 0) -> np.ndarray:
    rs = np.random.default_rng(1000 * seed + mode)
    if mode in (0, 1, 2):
        v = rd.random_unit_upper(rs, n) if mode else _unit(rs, n)
        return np.column_stack([v, rs.uniform(1e-4, 1, n).astype(F)])
    if mode == 3:
        return np.column_stack([rs.uniform(0, 1, n), rs.choice([1 / 1.5, 1.5, 1.33, 1 / 1.33, 0.001, 0.999], n)]).astype(F)
    if mode == 4:
        return rs.uniform(-0.1, 1.1, (n, 1)).astype(F)
    if mode in (5, 6):
        V, L = rd.random_unit_upper(rs, n), rd.random_unit_upper(rs, n)
        Fr = rs.uniform(0, 1, (n, 3)).astype(F)
        a = rs.uniform(1e-4, 1, n).astype(F)
        if mode == 5:
            L[::7, 2] *= -1  # early-out branch
            return np.column_stack([V, L, Fr, a])
        L[:, 2] *= -1
        L[::7, 2] *= -1
        return np.column_stack([V, L, Fr, a, rs.choice([1 / 1.5, 1.5, 1.33, 1 / 1.33], n).astype(F)])
    if mode == 7:
        return np.column_stack([rs.uniform(0, 1, (n, 2)).astype(F), rd.random_unit_upper(rs, n), rs.uniform(1e-4, 1, n).astype(F)])
    if mode == 8:
        return rs.uniform(0, 1, (n, 3)).astype(F)
    if mode == 9:
        L = rd.random_unit_upper(rs, n)
        L[::2, 2] *= -1
        return np.column_stack([rd.random_materials(rs, n), rd.random_unit_upper(rs, n), L])
    if mode == 10:
        seeds = rs.integers(1, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
        return np.column_stack([rd.random_materials(rs, n), rd.random_unit_upper(rs, n), _bits(seeds)])
    if mode == 11:
        a = rs.integers(0, 4096, (n, 4)).astype(np.uint32)
        a[:, 2] = rs.choice([512, 1920, 3840], n)
        return _bits(a)
    if mode == 12:
        rec = np.zeros((n, 42), F)
        rec[:, 0] = _bits(rs.integers(0, 1920, n))
        rec[:, 1] = _bits(rs.integers(0, 1080, n))
        rec[:, 2] = _bits(np.full(n, 1920))
        rec[:, 3] = _bits(np.full(n, 1080))
        rec[:, 4:8] = rs.uniform(0, 1, (n, 4))
        rec[:, 8] = np.where(rs.uniform(size=n) < 0.5, 0.0, 0.05)
        rec[:, 9] = 4.0
        vi = np.array([0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, 0, 3, 1, 0, 1], F)
        pi = np.array([0.414214, 0, 0, 0, 0, 0.414214, 0, 0, 0, 0, 0, 9.99, 0, 0, 1, 0.01], F)
        rec[:, 10:26] = vi + rs.normal(0, 0.05, (n, 16))
        rec[:, 26:42] = pi
        rec[::5, 26:42] += rs.normal(0, 0.05, (len(rec[::5]), 16))  # dense matrices: the mat4 * vec4 summation order matters
        return rec
    if mode == 13:
        o = rs.uniform(-10, 10, (n, 3)).astype(F)
        o[: n // 4] *= 1e-3
        return np.column_stack([o, _unit(rs, n)])
    if mode == 14:
        u = rs.uniform(0, 1, (n, 2)).astype(F)
        u[0] = 0.5  # offset == 0 branch
        return u
    if mode == 15:
        return _unit(rs, n)
    if mode == 16:
        v = [np.column_stack([_vec(rs, n, 2), rs.uniform(0, 1, (n, 2)).astype(F), _unit(rs, n)]) for _ in range(3)]
        v[1][::10, 3:5] = v[0][::10, 3:5]  # degenerate uv mapping: falls back to the vertex tangent frame
        v[2][::10, 3:5] = v[0][::10, 3:5]
        return np.column_stack(v + [_unit(rs, n), _unit(rs, n)])
    if mode == 17:
        return np.column_stack([_vec(rs, n, 2), _vec(rs, n, 2), _unit(rs, n), _vec(rs, n, 2), _unit(rs, n), _unit(rs, n)])
    if mode == 18:
        return np.column_stack([_vec(rs, n, 0.01), _vec(rs, n, 0.01), _vec(rs, n), _vec(rs, n)])
    if mode in (19, 20):
        der = rs.normal(0, 0.01, (n, 4)).astype(F)
        head = [der, _unit(rs, n), _vec(rs, n, 2), _unit(rs, n), _unit(rs, n), _vec(rs, n, 0.1), _vec(rs, n, 0.1)]
        tail = [_vec(rs, n, 2), _unit(rs, n), _vec(rs, n, 2), _unit(rs, n)]
        mid = [rs.uniform(0.5, 2, (n, 1)).astype(F)] if mode == 20 else []
        return np.column_stack(head + mid + tail)
    if mode == 21:
        bary = rs.dirichlet([1, 1, 1], n).astype(F)
        refr = (rs.uniform(size=(n, 1)) < 0.5).astype(F)
        return np.column_stack([_vec(rs, n, 2), _vec(rs, n, 2), _unit(rs, n), _vec(rs, n, 2), _unit(rs, n), _vec(rs, n, 2),
                                _unit(rs, n), bary, refr])
    if mode == 22:
        count = (rs.uniform(size=n) < 0.7).astype(np.uint32)
        return np.column_stack([rs.uniform(0, 1, (n, 3)).astype(F), _vec(rs, n, 3), rs.uniform(0, 10, (n, 3)).astype(F),
                                _vec(rs, n), rs.uniform(0, 10, (n, 3)).astype(F), _vec(rs, n, 3),
                                rs.uniform(0.1, 1, (n, 3)).astype(F), _bits(count)])
    if mode == 23:
        return np.column_stack([_vec(rs, n, 2), _unit(rs, n), _unit(rs, n), _unit(rs, n), _rows34(rs, n, 0.3), _rows34(rs, n, 0.3)])
    if mode == 24:
        return rs.uniform(0, 1, (n, 3)).astype(F)
    if mode == 25:
        return rs.uniform(0, 20, (n, 3)).astype(F)
    raise ValueError(mode)


def bit_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Per-record bitwise equality of two float32 arrays (NaNs of any payload count as equal)."""
    a, b = np.ascontiguousarray(a, F), np.ascontiguousarray(b, F)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return same.reshape(len(a), -1).all(axis=1)
