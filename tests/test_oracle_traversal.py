"""The oracle's own BVH + fp32 watertight triangle test against an fp64 brute force over all
triangles (no BVH) — the ray/box and ray/triangle stages have no source in the reference
(driver / RT hardware), so this is what anchors them (parity unpinned)."""
import numpy as np

import conftest

sc = conftest.pkg.scene


def random_rays(rs, n, lo, hi):
    rays = np.zeros(n, sc.RAY)
    rays["origin"] = rs.uniform(lo, hi, (n, 3))
    d = rs.normal(size=(n, 3))
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmin"] = 1e-5
    rays["tmax"] = 1e4
    return rays


def test_bvh_vs_bruteforce_f64(default_oracle):
    rs = np.random.default_rng(5)
    rays = random_rays(rs, 20000, [-7, -1.2, -2.2], [-2, 3.2, 2.2])
    a = default_oracle.trace_closest(rays)
    b, t64 = default_oracle.trace_closest_bruteforce_f64(rays)
    hit = b["instance"] != sc.NO_HIT
    assert hit.mean() > 0.5  # the box is open towards the camera
    same = (a["instance"] == b["instance"]) & (a["geometry"] == b["geometry"]) & (a["primitive"] == b["primitive"])
    # rays grazing an edge (a barycentric within 1e-4 of 0) may legitimately resolve differently
    w = 1 - b["u"] - b["v"]
    grazing = (np.minimum(np.minimum(b["u"], b["v"]), w) < 1e-4) & hit
    assert (same | grazing).all(), f"{(~(same | grazing)).sum()} non-grazing mismatches"
    ok = same & hit
    # 1e-5 relative, plus the fp32 resolution of the coordinates themselves for rays that start
    # within a hair of a surface (t ~ 1e-3 at |p| ~ 7: one ulp of p is already 1e-4 of t)
    err = np.abs(a["t"][ok].astype(np.float64) - t64[ok])
    scale = np.abs(rays["origin"][ok]).max(axis=1)
    assert (err <= 1e-5 * t64[ok] + 4 * np.finfo(np.float32).eps * scale).all()


def test_first_hit_aov_matches_trace(default_oracle, default_scene):
    p = default_scene.default_params()
    aov = default_oracle.first_hit_aov(p, 64, 64)
    assert (aov["instance"] != sc.NO_HIT).mean() > 0.8
    assert aov["t"][aov["instance"] != sc.NO_HIT].min() > 0


def test_occlusion_consistent_with_closest(default_oracle):
    rs = np.random.default_rng(6)
    rays = random_rays(rs, 5000, [-6.4, -0.9, -1.9], [-2.6, 2.9, 1.9])
    rays["tmax"] = rs.uniform(0.1, 6, len(rays))
    occ = default_oracle.trace_occlusion(rays)
    clo = default_oracle.trace_closest(rays)
    assert ((clo["instance"] != sc.NO_HIT) == (occ != 0)).all()  # all geometry is opaque here
