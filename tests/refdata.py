"""Input grids of the reference's own shader unit tests, regenerated from their definitions:
Path-Tracing-Tests/TestData.h:14-101 (edge-case values and the three generators),
ShadingTest.cpp:13-251 and BsdfTest.cpp:12-41 (how each test combines them)."""
import numpy as np

F = np.float32


def _normalize(v):
    v = np.asarray(v, F)
    return (v * (F(1) / np.sqrt(np.dot(v, v), dtype=F))).astype(F)


EDGE_VEC3S = [_normalize([0.99, 0.0, 0.01]), _normalize([0.0, 0.99, 0.01]), _normalize([0.01, 0.0, 0.99])]
EDGE_FLOATS = [F(0.001), F(0.999)]


def vec3_float():  # Vec3FloatGenerator: vec3 index runs fastest
    return [(v, f) for f in EDGE_FLOATS for v in EDGE_VEC3S]


def float_float():  # FloatFloatGenerator: first float runs fastest
    return [(a, b) for b in EDGE_FLOATS for a in EDGE_FLOATS]


def vec3_vec3():  # Vec3Vec3Generator: first vec3 runs fastest
    return [(a, b) for b in EDGE_VEC3S for a in EDGE_VEC3S]


def grid_vec3_float():
    """GGXDistribution / Lambda / GGXSmith inputs (ShadingTest.cpp:13-88): 6 x (xyz, alpha)."""
    return np.array([[*v, f] for v, f in vec3_float()], F)


def grid_dielectric():
    """ShadingTest.cpp:90-114: 4 x (VdotH, eta)."""
    return np.array([[a, b] for a, b in float_float()], F)


def grid_schlick():
    """ShadingTest.cpp:116-137."""
    return np.array([[f] for f in EDGE_FLOATS], F)


def grid_reflection():
    """ShadingTest.cpp:139-174: 9 (V, L) x 6 (F, alpha) = 54 x (V, L, F, alpha)."""
    return np.array([[*v, *l, *f, a] for v, l in vec3_vec3() for f, a in vec3_float()], F)


def grid_refraction(flip_l=False):
    """ShadingTest.cpp:176-216: 2 eta x 54 = 108 x (V, L, F, alpha, eta).  All reference inputs
    have L.z > 0 (the function early-outs); flip_l=True mirrors L below the surface so that the
    refraction body is exercised as well."""
    rows = []
    for eta in EDGE_FLOATS:
        for v, l in vec3_vec3():
            for f, a in vec3_float():
                ll = np.array([l[0], l[1], -l[2]], F) if flip_l else l
                rows.append([*v, *ll, *f, a, eta])
    return np.array(rows, F)


def grid_sample_ggx():
    """ShadingTest.cpp:218-251: 4 u x 6 (V, alpha) = 24 x (u.xy, V, alpha)."""
    return np.array([[u0, u1, *v, a] for u0, u1 in float_float() for v, a in vec3_float()], F)


def grid_lobe_pdfs():
    """BsdfTest.cpp:12-41: 125 x (Metalness, Transmission, F), index k*25 + i*5 + j."""
    fl = [0.0, 0.25, 0.5, 0.75, 1.0]
    return np.array([[fl[i], fl[j], fl[k]] for k in range(5) for i in range(5) for j in range(5)], F)


def random_materials(rng, n):
    """MaterialSample records (17 floats) for the evaluateBSDF / sampleBSDF parity tests."""
    m = np.zeros((n, 17), F)
    m[:, 0:3] = rng.uniform(0, 2, (n, 3))  # EmissiveColor
    m[:, 3:6] = rng.uniform(0.05, 1, (n, 3))  # Color
    m[:, 6:9] = [0, 0, 1]  # Normal (unused by the BSDF)
    m[:, 9] = np.maximum(rng.uniform(0, 1, n), 0.01)  # Roughness (already regularised)
    m[:, 10] = rng.choice([0.0, 0.3, 1.0], n) * rng.uniform(0.5, 1, n).round()  # Metalness
    m[:, 11] = rng.choice([0.0, 0.5, 1.0], n)  # Transmission
    ior = rng.uniform(1.1, 2.0, n)
    m[:, 12] = np.where(rng.uniform(0, 1, n) < 0.5, ior, 1 / ior)  # Eta
    m[:, 13:16] = rng.uniform(0.2, 1, (n, 3))
    m[:, 16] = rng.uniform(0.1, 10, n)
    return m


def random_unit_upper(rng, n, zmin=0.02):
    v = rng.normal(size=(n, 3))
    v[:, 2] = np.abs(v[:, 2]) + zmin
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(F)
