"""The reference's debug pipeline (Debug/debugRaygen.rgen, debugClosestHit.rchit, debugAnyhit.rahit, debugMiss.rmiss):
oracle restatement against independent closed forms, CUDA (pt_debug_render) against the oracle."""
import importlib

import numpy as np
import pytest

import conftest

core = conftest.core
sc = conftest.pkg.scene
scenes = importlib.import_module("path-tracing_b200.scenes")
W, H = 128, 96
FLAG = {"no_color": 0x01, "no_normal": 0x02, "no_mips": 0x04, "no_shadows": 0x08, "dx": 0x10}


def random_color(ids):
    """debugClosestHit.rchit:141-161 in numpy."""
    x = ids.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x = (x * np.uint64(0x1ECA7D79)) & m
    x ^= x >> np.uint64(20)
    x = ((x << np.uint64(8)) | (x >> np.uint64(24))) & m
    x = (~x) & m
    x ^= (x << np.uint64(5)) & m
    x = (x + np.uint64(0x10AFE4E7)) & m
    return np.stack([(x >> np.uint64(24)) & np.uint64(255), (x >> np.uint64(16)) & np.uint64(255), (x >> np.uint64(8)) & np.uint64(255)],
                    -1).astype(np.float32) / np.float32(255.0)


@pytest.fixture(scope="module")
def feature_oracle(oracle_mod):
    s = scenes.feature_scene()
    return s, oracle_mod.OracleScene(s)


def test_oracle_id_and_position_modes(feature_oracle):
    s, o = feature_oracle
    p = s.default_params()
    aov = o.first_hit_aov(p, W, H)
    hit = aov["instance"] != 0xFFFFFFFF
    for mode, field in (("instance", "instance"), ("geometry", "geometry"), ("primitive", "primitive")):
        img = o.debug_render(p, W, H, mode)
        assert np.array_equal(img[hit][:, :3], random_color(aov[field][hit])), mode
        assert (img[~hit][:, :3] == np.float32(0.2)).all() and (img[..., 3] == 1).all()
    # world position = origin + t * direction of the pixel-centre ray
    pos = o.debug_render(p, W, H, "world_position")[..., :3]
    vi = np.asarray(s.view_inverse, np.float64).reshape(4, 4).T
    pi = np.asarray(s.proj_inverse, np.float64).reshape(4, 4).T
    ys, xs = np.nonzero(hit)
    d = np.stack([2 * (xs + 0.5) / W - 1, 2 * (ys + 0.5) / H - 1, np.ones(len(xs)), np.ones(len(xs))], 1)
    t = (pi @ d.T).T
    dirv = t[:, :3] / np.linalg.norm(t[:, :3], axis=1, keepdims=True)
    want = vi[:3, 3] + (vi[:3, :3] @ dirv.T).T * aov["t"][ys, xs][:, None]
    assert np.allclose(pos[ys, xs], want, rtol=1e-4, atol=1e-4)
    n = o.debug_render(p, W, H, "normal")[..., :3]
    assert np.allclose(np.linalg.norm(n[hit], axis=1), 1.0, atol=1e-5)
    uv = o.debug_render(p, W, H, "texture_coords")
    assert (uv[..., 2][hit] == 0).all()


def test_oracle_flags(feature_oracle):
    s, o = feature_oracle
    p = s.default_params()
    base = o.debug_render(p, W, H, "color")
    assert np.isfinite(base).all() and base[..., :3].max() > 0.3
    # without shadow rays the frame gets brighter (not every pixel: specular-glossiness materials can have a
    # "metalness" outside [0, 1] and with it a negative diffuse weight, material.glsl:108-110)
    lit = o.debug_render(p, W, H, "color", hit_group_flags=FLAG["no_shadows"])
    assert lit[..., :3].mean() > base[..., :3].mean() and (lit[..., :3] > base[..., :3] + 1e-3).any()
    assert ((lit[..., :3] >= base[..., :3] - 1e-6).all(-1)).mean() > 0.9
    # mips disabled: lod 0 everywhere -> the mips view is exactly 1
    mips = o.debug_render(p, W, H, "mips", hit_group_flags=FLAG["no_mips"])
    hit = o.first_hit_aov(p, W, H)["instance"] != 0xFFFFFFFF
    assert (mips[hit][:, 0] == 1.0).all()
    # the normal texture changes the normal view, disabling it restores the interpolated vertex normal (+ the default texel)
    a, b = o.debug_render(p, W, H, "normal"), o.debug_render(p, W, H, "normal", hit_group_flags=FLAG["no_normal"])
    assert not np.array_equal(a, b)
    # force-opaque: the alpha-tested leaf cards hide what is behind them
    c = o.debug_render(p, W, H, "primitive", raygen_flags=0x1)
    assert not np.array_equal(c, o.debug_render(p, W, H, "primitive"))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", core.DEBUG_MODES)
def test_gpu_matches_oracle(feature_oracle, mode):
    s, o = feature_oracle
    p = s.default_params()
    with core.Renderer(0) as r:
        r.update_scene_data(s)
        for raygen, hitflags in ((0, 0), (0x1, 0), (0x2, 0), (0x3, 0), (0, FLAG["no_color"] | FLAG["no_normal"]),
                                 (0, FLAG["no_mips"] | FLAG["no_shadows"]), (0, FLAG["dx"])):
            got = r.debug_render(p, W, H, mode, raygen, hitflags)
            want = o.debug_render(p, W, H, mode, raygen, hitflags)
            if mode in ("geometry", "primitive", "instance"):
                assert np.array_equal(got, want), (mode, raygen, hitflags)
            else:
                close = np.isclose(got, want, rtol=2e-4, atol=2e-5).all(-1)
                assert close.mean() > 0.995, (mode, raygen, hitflags, close.mean())


@pytest.mark.gpu
def test_gpu_default_scene_and_sky(default_renderer, default_oracle, default_scene):
    p = default_scene.default_params()
    for mode in ("color", "normal", "mips"):
        got = default_renderer.debug_render(p, 160, 160, mode)
        want = default_oracle.debug_render(p, 160, 160, mode)
        assert np.isclose(got, want, rtol=2e-4, atol=2e-5).all(-1).mean() > 0.995, mode
    # back-face culling (gl_RayFlagsCullBackFacingTrianglesEXT): ids bit for bit with the oracle
    got = default_renderer.debug_render(p, 160, 160, "primitive", raygen_flags=0x2)
    assert np.array_equal(got, default_oracle.debug_render(p, 160, 160, "primitive", raygen_flags=0x2))


def test_oracle_back_face_culling(oracle_mod):
    """gl_RayFlagsCullBackFacingTrianglesEXT against a closed form: a single quad seen from its front, from its back,
    and through a MIRRORING instance transform (negative determinant), which reverses the world-space winding but not
    the object-space facing Vulkan culls by.  Front = the vertices appear clockwise from the ray origin."""
    def scene_with(instance_matrix, eye):
        b = scenes.SceneBuilder()
        m = b.add_material_mr(color=(0.8, 0.8, 0.8, 1))
        g = b.add_geometry(*scenes.quad(2.0, 2.0))  # in the xz-plane; triangles (0, 1, 2), (2, 3, 0)
        b.add_instance(b.add_model([(g, m, None)]), instance_matrix)
        eye = np.asarray(eye, np.float64)
        return b.build(scenes.camera_matrices(eye, -eye, 32, 32, fov_deg=40, up=(0.0, 0.0, 1.0)), (32, 32))

    quad_v, quad_i = scenes.quad(2.0, 2.0)
    p0, p1, p2 = (quad_v["position"][quad_i[k]].astype(np.float64) for k in range(3))
    n_obj = np.cross(p1 - p0, p2 - p0)  # numeric normal of the object-space winding
    ident = np.eye(4)
    mirror = np.diag([1.0, -1.0, 1.0, 1.0])  # flips y: the quad stays in place, the determinant is negative
    for matrix, mirrored in ((ident, False), (mirror, True)):
        for eye in ((0.0, 3.0, 0.0), (0.0, -3.0, 0.0)):
            s = scene_with(matrix, eye)
            o = oracle_mod.OracleScene(s)
            p = s.default_params()
            plain = o.debug_render(p, 32, 32, "instance")
            culled = o.debug_render(p, 32, 32, "instance", raygen_flags=0x2)
            centre_hit = not np.allclose(plain[16, 16, :3], 0.2)
            assert centre_hit
            # object-space ray direction at the centre: world direction is -eye; the mirror maps it back
            d_obj = -np.asarray(eye) * (np.array([1.0, -1.0, 1.0]) if mirrored else 1.0)
            front = np.dot(n_obj, d_obj) < 0
            still_hit = not np.allclose(culled[16, 16, :3], 0.2)
            assert still_hit == front, (mirrored, eye)
