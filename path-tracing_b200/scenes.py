"""Procedural scenes, built through a Python mirror of the reference's ``SceneBuilder``
(Path-Tracing/Scene.h:268-361): the same calls (AddGeometry / AddMaterial / AddTexture / AddModel /
AddModelInstance / AddLight ...) produce the same flat arrays the reference's Scene holds, so a
generated scene crosses the C ABI exactly like one loaded by the reference's importers.

Scenes:
  * ``feature_scene``  — small scene touching every hot-path feature (mesh + instance transforms,
    three material models, alpha-tested cards, transmission + attenuation, point + directional
    lights, textures with mips).  Parity-test fodder.
  * ``chess_scene``    — BASELINE.json configs[1]: "ABeautifulGame-class" stand-in (the Khronos
    asset is not available offline): a chessboard with 32 lathe-turned pieces.  Triangle and
    material counts are DECLARED BY THIS BUILDER (see ``chess_scene.__doc__``), not quoted from
    the real asset.
All content is generated from fixed integer seeds.
"""
from __future__ import annotations

import math

import os

import numpy as np

from . import scene as sc

F = np.float32


# ---------------------------------------------------------------------------------------------
# camera: glm with GLM_FORCE_LEFT_HANDED + GLM_FORCE_DEPTH_ZERO_TO_ONE (Core/Camera.cpp:1-2,52-71)
# ---------------------------------------------------------------------------------------------
def look_at_lh(eye, center, up) -> np.ndarray:
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(up, f)
    s /= np.linalg.norm(s)
    u = np.cross(f, s)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, -f @ eye
    return m


def perspective_fov_lh_zo(fov_deg, width, height, near, far) -> np.ndarray:
    h = 1.0 / math.tan(math.radians(fov_deg) / 2)
    w = h * height / width
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = w, h
    m[2, 2] = far / (far - near)
    m[2, 3] = -(far * near) / (far - near)
    m[3, 2] = 1.0
    return m


def camera_matrices(position, direction, width, height, fov_deg=45.0, near=100.0, far=0.1, up=(0.0, -1.0, 0.0)):
    """(ViewInverse, ProjInverse) as 16 floats, column-major like glm.  Defaults are the reference's
    InputCamera (Scene.h:259-260 — note its near/far arguments are 100 / 0.1)."""
    position = np.asarray(position, np.float64)
    view = look_at_lh(position, position + np.asarray(direction, np.float64), up)
    proj = perspective_fov_lh_zo(fov_deg, width, height, near, far)
    return (np.linalg.inv(view).T.astype(F).reshape(-1), np.linalg.inv(proj).T.astype(F).reshape(-1))


# ---------------------------------------------------------------------------------------------
# SceneBuilder mirror
# ---------------------------------------------------------------------------------------------
def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def scale(x, y=None, z=None):
    y = x if y is None else y
    z = x if z is None else z
    return np.diag([x, y, z, 1.0])


def rotate_y(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def rotate_x(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    m = np.eye(4)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


class SceneBuilder:
    """Python counterpart of PathTracing::SceneBuilder.  Matrices are standard column-vector 4x4
    (world = M @ [p, 1]); the 3x4 row-major form the reference stores is produced at build()."""

    def __init__(self):
        self.vertices = []
        self.indices = []
        self.vertex_count = 0
        self.index_count = 0
        self.transforms = [np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], F)]  # IdentityTransformIndex = 0
        self.geometries = []
        self.mesh_records = []
        self.models = []
        self.instances = []
        self.mr, self.sg, self.phong = [], [], []
        self.textures = []
        self.point_lights = []
        self.directional = np.zeros((), sc.DIRECTIONAL_LIGHT)
        # SceneBuilder::g_DefaultLight (Scene.h:352-355)
        self.directional["color"] = (10, 10, 10)
        self.directional["direction"] = (-0.4, -1.0, -0.2)
        self.skybox = None
        self.hit_flags = sc.HIT_FLAGS_NONE
        self.animated_vertices, self.animated_indices, self.animated_flags = [], [], {}

    def add_geometry(self, vertices: np.ndarray, indices: np.ndarray, is_opaque: bool = True) -> int:
        vertices = np.ascontiguousarray(vertices, sc.VERTEX)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        assert len(indices) % 3 == 0 and (len(indices) == 0 or indices.max() < len(vertices))
        self.geometries.append((self.vertex_count, len(vertices), self.index_count, len(indices), 1 if is_opaque else 0))
        self.vertices.append(vertices)
        self.indices.append(indices)
        self.vertex_count += len(vertices)
        self.index_count += len(indices)
        return len(self.geometries) - 1

    def add_animated_geometry(self, vertices: np.ndarray, indices: np.ndarray, bone_indices: np.ndarray,
                              bone_weights: np.ndarray, is_opaque: bool = True) -> int:
        """A skinned geometry (Geometry::IsAnimated): its offsets address the animated vertex / index buffers
        (SceneBuilder::AddGeometry's animated overload, Scene.cpp:100-123)."""
        vertices = np.ascontiguousarray(vertices, sc.VERTEX)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        assert len(indices) % 3 == 0 and indices.max() < len(vertices)
        av = np.zeros(len(vertices), sc.ANIMATED_VERTEX)
        for f in ("position", "texcoords", "normal", "tangent", "bitangent"):
            av[f] = vertices[f]
        av["bone_indices"], av["bone_weights"] = bone_indices, bone_weights
        a_v = sum(len(v) for v in self.animated_vertices)
        a_i = sum(len(i) for i in self.animated_indices)
        self.geometries.append((a_v, len(vertices), a_i, len(indices), 1 if is_opaque else 0))
        self.animated_flags[len(self.geometries) - 1] = 1
        self.animated_vertices.append(av)
        self.animated_indices.append(indices)
        return len(self.geometries) - 1

    def add_texture(self, pixels: np.ndarray, srgb: bool) -> int:
        """Returns the bindless slot, SceneTextureOffset + i (Scene.cpp:125-140)."""
        self.textures.append(sc.Texture(np.ascontiguousarray(pixels), srgb))
        return sc.SCENE_TEXTURE_OFFSET + len(self.textures) - 1

    def add_material_mr(self, color=(1, 1, 1, 1), roughness=1.0, metalness=0.0, ior=1.5, transmission=0.0,
                        emissive=(0, 0, 0), emissive_intensity=0.0, attenuation_color=(1, 1, 1), attenuation_distance=1e32,
                        color_idx=sc.TEX_DEFAULT_COLOR, normal_idx=sc.TEX_DEFAULT_NORMAL, roughness_idx=sc.TEX_DEFAULT_ROUGHNESS,
                        metallic_idx=sc.TEX_DEFAULT_METALLIC, emissive_idx=sc.TEX_DEFAULT_EMISSIVE) -> int:
        m = np.zeros((), sc.MATERIAL_MR)
        m["emissive_color"], m["emissive_intensity"] = emissive, emissive_intensity
        m["color"], m["roughness"], m["metalness"], m["ior"], m["transmission"] = color, roughness, metalness, ior, transmission
        m["attenuation_color"], m["attenuation_distance"] = attenuation_color, attenuation_distance
        m["emissive_idx"], m["color_idx"], m["normal_idx"] = emissive_idx, color_idx, normal_idx
        m["roughness_idx"], m["metallic_idx"] = roughness_idx, metallic_idx
        self.mr.append(m)
        return sc.material_id(len(self.mr) - 1, sc.MATERIAL_TYPE_MR)

    def _add_sg(self, store, mtype, color, specular, gloss, ior, transmission, emissive, emissive_intensity,
                attenuation_color, attenuation_distance, color_idx, normal_idx, specular_idx, gloss_idx, emissive_idx):
        m = np.zeros((), sc.MATERIAL_SG)
        m["emissive_color"], m["emissive_intensity"] = emissive, emissive_intensity
        m["color"], m["specular"], m["glossiness"] = color, specular, gloss
        m["attenuation_color"], m["attenuation_distance"] = attenuation_color, attenuation_distance
        m["ior"], m["transmission"] = ior, transmission
        m["emissive_idx"], m["color_idx"], m["normal_idx"] = emissive_idx, color_idx, normal_idx
        m["specular_idx"], m["glossiness_idx"] = specular_idx, gloss_idx
        store.append(m)
        return sc.material_id(len(store) - 1, mtype)

    def add_material_sg(self, color=(1, 1, 1, 1), specular=(1, 1, 1), glossiness=1.0, ior=1.5, transmission=0.0,
                        emissive=(0, 0, 0), emissive_intensity=0.0, attenuation_color=(1, 1, 1), attenuation_distance=1e32,
                        color_idx=sc.TEX_DEFAULT_COLOR, normal_idx=sc.TEX_DEFAULT_NORMAL, specular_idx=sc.TEX_DEFAULT_SPECULAR,
                        glossiness_idx=sc.TEX_DEFAULT_GLOSSINESS, emissive_idx=sc.TEX_DEFAULT_EMISSIVE) -> int:
        return self._add_sg(self.sg, sc.MATERIAL_TYPE_SG, color, specular, glossiness, ior, transmission, emissive,
                            emissive_intensity, attenuation_color, attenuation_distance, color_idx, normal_idx, specular_idx,
                            glossiness_idx, emissive_idx)

    def add_material_phong(self, color=(1, 1, 1, 1), specular=(1, 1, 1), shininess=1.0, ior=1.5, transmission=0.0,
                           emissive=(0, 0, 0), emissive_intensity=0.0, attenuation_color=(1, 1, 1), attenuation_distance=1e32,
                           color_idx=sc.TEX_DEFAULT_COLOR, normal_idx=sc.TEX_DEFAULT_NORMAL, specular_idx=sc.TEX_DEFAULT_SPECULAR,
                           shininess_idx=sc.TEX_DEFAULT_SHININESS, emissive_idx=sc.TEX_DEFAULT_EMISSIVE) -> int:
        return self._add_sg(self.phong, sc.MATERIAL_TYPE_PHONG, color, specular, shininess, ior, transmission, emissive,
                            emissive_intensity, attenuation_color, attenuation_distance, color_idx, normal_idx, specular_idx,
                            shininess_idx, emissive_idx)

    def add_model(self, meshes) -> int:
        """meshes: iterable of (geometry_index, material_id, transform 4x4 or None) — MeshInfo, Scene.h:80-86."""
        offset = len(self.mesh_records)
        count = 0
        for geometry, material, transform in meshes:
            tindex = 0
            if transform is not None and not np.allclose(transform, np.eye(4)):
                self.transforms.append(np.asarray(transform, np.float64)[:3, :].astype(F).reshape(-1))
                tindex = len(self.transforms) - 1
            self.mesh_records.append((geometry, material, tindex))
            count += 1
        self.models.append((offset, count))
        return len(self.models) - 1

    def add_instance(self, model: int, transform=None) -> int:
        t = np.eye(4) if transform is None else np.asarray(transform, np.float64)
        self.instances.append((t[:3, :].astype(F).reshape(-1), model))
        return len(self.instances) - 1

    def add_light(self, color, position, constant=1.0, linear=0.0, quadratic=1.0):
        l = np.zeros((), sc.POINT_LIGHT)
        l["color"], l["position"] = color, position
        l["attenuation_constant"], l["attenuation_linear"], l["attenuation_quadratic"] = constant, linear, quadratic
        self.point_lights.append(l)

    def set_directional_light(self, color, direction):
        self.directional["color"], self.directional["direction"] = color, direction

    def set_skybox_2d(self, pixels: np.ndarray, srgb: bool = False):
        self.skybox = sc.Texture(np.ascontiguousarray(pixels), srgb)

    def set_skybox_cube(self, faces, srgb: bool = False):
        """faces: Front, Back, Up, Down, Left, Right (SkyboxCube, Scene.h:136-144)."""
        assert len(faces) == 6
        self.skybox = [sc.Texture(np.ascontiguousarray(f), srgb) for f in faces]

    def build(self, camera=None, extent=(0, 0)) -> sc.SceneData:
        s = sc.SceneData()
        s.vertices = np.concatenate(self.vertices) if self.vertices else np.zeros(0, sc.VERTEX)
        s.indices = np.concatenate(self.indices) if self.indices else np.zeros(0, np.uint32)
        s.transforms = np.stack(self.transforms).astype(F)
        s.geometries = np.array(self.geometries, sc.GEOMETRY) if self.geometries else np.zeros(0, sc.GEOMETRY)
        s.mesh_records = np.array(self.mesh_records, sc.MESH_RECORD) if self.mesh_records else np.zeros(0, sc.MESH_RECORD)
        s.models = np.array(self.models, sc.MODEL) if self.models else np.zeros(0, sc.MODEL)
        inst = np.zeros(len(self.instances), sc.INSTANCE)
        for i, (t, m) in enumerate(self.instances):
            inst[i]["transform"], inst[i]["model_index"] = t, m
        s.instances = inst
        s.mr_materials = np.array(self.mr, sc.MATERIAL_MR) if self.mr else np.zeros(0, sc.MATERIAL_MR)
        s.sg_materials = np.array(self.sg, sc.MATERIAL_SG) if self.sg else np.zeros(0, sc.MATERIAL_SG)
        s.phong_materials = np.array(self.phong, sc.MATERIAL_SG) if self.phong else np.zeros(0, sc.MATERIAL_SG)
        s.textures = list(self.textures)
        s.point_lights = np.array(self.point_lights, sc.POINT_LIGHT) if self.point_lights else np.zeros(0, sc.POINT_LIGHT)
        s.directional_light = self.directional.copy()
        if isinstance(self.skybox, list):
            s.skybox_cube = self.skybox
            s.miss_flags = sc.MISS_FLAGS_SKYBOX_CUBE
        else:
            s.skybox_2d = self.skybox
            s.miss_flags = sc.MISS_FLAGS_SKYBOX_2D if self.skybox is not None else sc.MISS_FLAGS_NONE
        s.hit_flags = self.hit_flags
        if self.animated_vertices:
            s.animated_vertices = np.concatenate(self.animated_vertices)
            s.animated_indices = np.concatenate(self.animated_indices)
            flags = np.zeros(len(self.geometries), np.uint32)
            flags[list(self.animated_flags)] = 1
            s.geometry_is_animated = flags
            s.bone_transforms = np.tile(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], F), (int(s.animated_vertices["bone_indices"].max()) + 1, 1))
        if camera is not None:
            s.view_inverse, s.proj_inverse = camera
            s.camera_extent = tuple(extent)
        return s


# ---------------------------------------------------------------------------------------------
# mesh primitives (vertices carry position, uv, normal, tangent, bitangent like Shaders::Vertex)
# ---------------------------------------------------------------------------------------------
def _vertices(pos, uv, nrm, tan, bit) -> np.ndarray:
    v = np.zeros(len(pos), sc.VERTEX)
    v["position"], v["texcoords"], v["normal"], v["tangent"], v["bitangent"] = pos, uv, nrm, tan, bit
    return v


def quad(size_x=1.0, size_z=1.0, uv_scale=1.0):
    """Horizontal quad in the xz-plane facing +y, two triangles."""
    x, z = size_x / 2, size_z / 2
    pos = np.array([[-x, 0, z], [x, 0, z], [x, 0, -z], [-x, 0, -z]], F)
    uv = np.array([[0, 1], [1, 1], [1, 0], [0, 0]], F) * uv_scale
    n = np.tile(np.array([0, 1, 0], F), (4, 1))
    t = np.tile(np.array([1, 0, 0], F), (4, 1))
    b = np.tile(np.array([0, 0, -1], F), (4, 1))
    return _vertices(pos, uv, n, t, b), np.array([0, 1, 2, 2, 3, 0], np.uint32)


def grid(nx, nz, size_x=1.0, size_z=1.0, uv_scale=1.0, height=None):
    """Tessellated horizontal patch; height(x, z) -> y optionally displaces it."""
    xs = np.linspace(-size_x / 2, size_x / 2, nx + 1)
    zs = np.linspace(-size_z / 2, size_z / 2, nz + 1)
    X, Z = np.meshgrid(xs, zs, indexing="xy")
    Y = np.zeros_like(X) if height is None else height(X, Z)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3)
    uv = np.stack([(X / size_x + 0.5) * uv_scale, (0.5 - Z / size_z) * uv_scale], -1).reshape(-1, 2)
    if height is None:
        n = np.tile([0.0, 1.0, 0.0], (len(pos), 1))
    else:
        e = 1e-3 * max(size_x, size_z)
        dx = (height(X + e, Z) - height(X - e, Z)) / (2 * e)
        dz = (height(X, Z + e) - height(X, Z - e)) / (2 * e)
        n = np.stack([-dx, np.ones_like(dx), -dz], -1).reshape(-1, 3)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
    t = np.cross(n, [0, 0, -1.0])
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    i = (np.arange(nz)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    idx = np.stack([i, i + nx + 1, i + nx + 2, i, i + nx + 2, i + 1], -1).reshape(-1)  # CCW seen from +y
    return _vertices(pos, uv, n, t, b), idx.astype(np.uint32)


def box(sx=1.0, sy=1.0, sz=1.0):
    """Axis-aligned box centred at the origin, 24 vertices / 12 triangles, per-face frames."""
    faces = [  # normal, tangent, bitangent
        ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (-1, 0, 0), (0, 1, 0)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
        ((1, 0, 0), (0, 0, -1), (0, 1, 0)), ((0, 1, 0), (1, 0, 0), (0, 0, -1)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
    ]
    half = np.array([sx, sy, sz]) / 2
    pos, uv, nrm, tan, bit, idx = [], [], [], [], [], []
    for k, (n, t, b) in enumerate(faces):
        n, t, b = (np.array(v, np.float64) for v in (n, t, b))
        for (a, c), (u, v) in zip(((-1, -1), (1, -1), (1, 1), (-1, 1)), ((0, 1), (1, 1), (1, 0), (0, 0))):
            pos.append((n + a * t + c * b) * half)
            uv.append((u, v))
            nrm.append(n), tan.append(t), bit.append(b)
        idx += [4 * k + i for i in (0, 1, 2, 2, 3, 0)]
    return _vertices(np.array(pos), np.array(uv), np.array(nrm), np.array(tan), np.array(bit)), np.array(idx, np.uint32)


def lathe(profile: np.ndarray, segments: int, uv_scale=(1.0, 1.0)):
    """Surface of revolution around +y of a profile [(radius, y), ...] (bottom to top)."""
    profile = np.asarray(profile, np.float64)
    rings = len(profile)
    d = np.gradient(profile, axis=0)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    theta = np.linspace(0, 2 * np.pi, segments + 1)
    c, s = np.cos(theta), np.sin(theta)
    r, y = profile[:, 0][:, None], profile[:, 1][:, None]
    pos = np.stack([r * c, np.broadcast_to(y, (rings, segments + 1)), r * s], -1)
    # outward normal of the profile curve: (dy, -dr) rotated around the axis
    nr, ny = d[:, 1][:, None], -d[:, 0][:, None]
    nrm = np.stack([nr * c, np.broadcast_to(ny, (rings, segments + 1)), nr * s], -1)
    tan = np.stack([-s, np.zeros_like(s), c], -1)[None].repeat(rings, 0)
    bit = np.stack([d[:, 0][:, None] * c, np.broadcast_to(d[:, 1][:, None], (rings, segments + 1)), d[:, 0][:, None] * s], -1)
    arc = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(profile, axis=0), axis=1))])
    arc /= max(arc[-1], 1e-12)
    uv = np.stack([np.broadcast_to(theta / (2 * np.pi) * uv_scale[0], (rings, segments + 1)),
                   np.broadcast_to((1 - arc)[:, None] * uv_scale[1], (rings, segments + 1))], -1)
    i = (np.arange(rings - 1)[:, None] * (segments + 1) + np.arange(segments)[None, :]).reshape(-1)
    idx = np.stack([i, i + segments + 1, i + segments + 2, i + segments + 2, i + 1, i], -1).reshape(-1)
    flat = lambda a: a.reshape(-1, a.shape[-1])
    return _vertices(flat(pos), flat(uv), flat(nrm), flat(tan), flat(bit)), idx.astype(np.uint32)


def sphere(radius=1.0, segments=32, rings=16):
    phi = np.linspace(-np.pi / 2, np.pi / 2, rings + 1)
    return lathe(np.stack([radius * np.cos(phi), radius * np.sin(phi)], -1), segments)


# ---------------------------------------------------------------------------------------------
# procedural textures
# ---------------------------------------------------------------------------------------------
def value_noise(rs: np.random.Generator, size: int, octaves: int = 5, base: int = 4) -> np.ndarray:
    """Tileable fractal value noise in [0, 1], (size, size) float32."""
    out = np.zeros((size, size), np.float64)
    amp, total = 1.0, 0.0
    for o in range(octaves):
        n = base << o
        if n > size:
            break
        lattice = rs.uniform(0, 1, (n, n))
        t = np.arange(size) * (n / size)
        i0 = np.floor(t).astype(int)
        f = t - i0
        f = f * f * (3 - 2 * f)
        i1 = (i0 + 1) % n
        rows = lattice[i0][:, i0] * (1 - f)[None, :] + lattice[i0][:, i1] * f[None, :]
        rows1 = lattice[i1][:, i0] * (1 - f)[None, :] + lattice[i1][:, i1] * f[None, :]
        out += amp * (rows * (1 - f)[:, None] + rows1 * f[:, None])
        total += amp
        amp *= 0.5
    return (out / total).astype(F)


def rgba8(r, g, b, a=None) -> np.ndarray:
    a = np.ones_like(r) if a is None else a
    return (np.clip(np.stack([r, g, b, a], -1), 0, 1) * 255 + 0.5).astype(np.uint8)


def normal_map_from_height(h: np.ndarray, strength: float) -> np.ndarray:
    dx = (np.roll(h, -1, 1) - np.roll(h, 1, 1)) * strength
    dy = (np.roll(h, -1, 0) - np.roll(h, 1, 0)) * strength
    n = np.stack([-dx, -dy, np.ones_like(h)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    return rgba8(n[..., 0] * 0.5 + 0.5, n[..., 1] * 0.5 + 0.5, n[..., 2] * 0.5 + 0.5)


# ---------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------
def feature_scene(seed: int = 1234, width: int = 160, height: int = 120, texture_size: int = 64) -> sc.SceneData:
    """A few hundred triangles exercising every branch of the hot path (see module docstring)."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size
    noise = value_noise(rs, n, 4)
    checker = ((np.indices((n, n)).sum(0) // (n // 8)) % 2).astype(F)
    color_tex = b.add_texture(rgba8(0.2 + 0.6 * noise, 0.3 + 0.5 * checker, 0.8 - 0.5 * noise), srgb=True)
    normal_tex = b.add_texture(normal_map_from_height(noise, 4.0), srgb=False)
    orm_tex = b.add_texture(rgba8(np.ones_like(noise), 0.2 + 0.8 * noise, checker), srgb=False)
    # leaf card: binary alpha from thresholded noise, plus a soft band of 0 < a < 1 and RGB zeroed where a == 0
    # like the reference's importer does (TextureImporter.cpp:24-51)
    alpha = np.clip((noise - 0.45) * 8, 0, 1)
    leaf = rgba8(0.1 + 0.2 * noise, 0.5 + 0.4 * noise, 0.1 * np.ones_like(noise), alpha)
    leaf[leaf[..., 3] == 0, :3] = 0
    leaf_tex = b.add_texture(leaf, srgb=True)
    emissive_tex = b.add_texture(rgba8(checker, 0.5 * checker, 0.1 * checker), srgb=True)
    spec_tex = b.add_texture(rgba8(0.3 + 0.5 * noise, 0.3 + 0.5 * noise, 0.3 + 0.5 * noise, 0.3 + 0.6 * checker), srgb=True)

    m_floor = b.add_material_mr(color=(0.9, 0.9, 0.9, 1), roughness=0.8, color_idx=color_tex, normal_idx=normal_tex,
                                roughness_idx=orm_tex, metallic_idx=orm_tex, metalness=0.6)
    m_red = b.add_material_mr(color=(0.8, 0.1, 0.1, 1), roughness=0.4)
    m_gold = b.add_material_mr(color=(1.0, 0.77, 0.34, 1), roughness=0.25, metalness=1.0)
    m_glass = b.add_material_mr(color=(0.9, 0.95, 1.0, 1), roughness=0.05, transmission=1.0, ior=1.5,
                                attenuation_color=(0.6, 0.9, 0.7), attenuation_distance=0.8)
    m_lamp = b.add_material_mr(color=(1, 1, 1, 1), emissive=(1.0, 0.9, 0.7), emissive_intensity=6.0, emissive_idx=emissive_tex)
    m_leaf = b.add_material_mr(color=(1, 1, 1, 1), roughness=0.7, color_idx=leaf_tex)
    m_sg = b.add_material_sg(color=(0.5, 0.5, 0.9, 1), specular=(0.9, 0.8, 0.7), glossiness=0.8, specular_idx=spec_tex,
                             glossiness_idx=spec_tex)
    m_phong = b.add_material_phong(color=(0.3, 0.8, 0.4, 1), specular=(0.5, 0.5, 0.5), shininess=0.6, shininess_idx=spec_tex)

    g_floor = b.add_geometry(*grid(6, 6, 8, 8, uv_scale=3.0, height=lambda x, z: 0.15 * np.sin(1.3 * x) * np.cos(0.9 * z)))
    g_box = b.add_geometry(*box(1, 1, 1))
    g_sphere = b.add_geometry(*sphere(0.5, 20, 10))
    g_card = b.add_geometry(*quad(1.5, 1.5), is_opaque=False)
    g_lamp = b.add_geometry(*quad(1.0, 1.0))
    vase_profile = [(0.25, 0.0), (0.4, 0.2), (0.3, 0.6), (0.15, 0.9), (0.25, 1.1)]
    g_vase = b.add_geometry(*lathe(vase_profile, 16))

    floor = b.add_model([(g_floor, m_floor, None)])
    # one model with two meshes, the second with a non-identity mesh transform (BLAS-baked transform)
    stack = b.add_model([(g_box, m_red, None), (g_sphere, m_gold, translate(0, 0.95, 0) @ scale(1.0, 0.8, 1.0))])
    glass = b.add_model([(g_sphere, m_glass, None)])
    vase = b.add_model([(g_vase, m_sg, None)])
    blob = b.add_model([(g_sphere, m_phong, scale(1.2, 0.6, 1.2))])
    cards = b.add_model([(g_card, m_leaf, rotate_x(90)), (g_card, m_leaf, translate(0.3, 0, 0.4) @ rotate_x(90) @ rotate_y(40))])
    lamp = b.add_model([(g_lamp, m_lamp, None)])

    b.add_instance(floor)
    b.add_instance(stack, translate(-1.6, 0.65, 0.8) @ rotate_y(30))
    b.add_instance(stack, translate(1.9, 0.45, 1.4) @ rotate_y(-50) @ scale(0.7, 0.7, 0.7))
    b.add_instance(glass, translate(0.2, 0.75, -0.3) @ scale(1.3, 1.3, 1.3))
    b.add_instance(vase, translate(-0.4, 0.1, 1.9))
    b.add_instance(blob, translate(1.4, 0.5, -1.2))
    b.add_instance(cards, translate(-1.0, 1.0, -1.6))
    b.add_instance(cards, translate(2.2, 1.0, -0.4) @ rotate_y(70) @ scale(0.8, 1.1, 0.8))
    b.add_instance(lamp, translate(0, 3.2, 0) @ rotate_x(180))

    b.add_light((6.0, 5.0, 4.0), (-2.5, 2.5, 2.0), 1.0, 0.1, 0.3)
    b.add_light((2.0, 3.0, 6.0), (2.5, 1.8, -2.0), 1.0, 0.0, 0.5)
    b.set_directional_light((2.5, 2.4, 2.2), (-0.4, -1.0, -0.2))
    eye = np.array([0.3, 2.2, 5.2])
    cam = camera_matrices(eye, np.array([0, 0.6, 0]) - eye, width, height, fov_deg=50)
    return b.build(cam, (width, height))


# chess-piece profiles: (radius, height) pairs, unit = one board square
_PIECES = {
    "pawn": [(0.0, 0.0), (0.30, 0.0), (0.32, 0.05), (0.26, 0.12), (0.14, 0.22), (0.11, 0.42), (0.17, 0.47), (0.11, 0.52),
             (0.16, 0.60), (0.18, 0.70), (0.13, 0.80), (0.0, 0.84)],
    "rook": [(0.0, 0.0), (0.34, 0.0), (0.36, 0.06), (0.28, 0.14), (0.20, 0.30), (0.19, 0.62), (0.27, 0.68), (0.28, 0.90),
             (0.20, 0.90), (0.20, 0.82), (0.0, 0.82)],
    "knight": [(0.0, 0.0), (0.34, 0.0), (0.36, 0.06), (0.27, 0.15), (0.18, 0.30), (0.22, 0.55), (0.27, 0.75), (0.20, 0.95),
               (0.10, 1.05), (0.0, 1.08)],
    "bishop": [(0.0, 0.0), (0.33, 0.0), (0.35, 0.06), (0.26, 0.15), (0.13, 0.32), (0.10, 0.68), (0.19, 0.74), (0.10, 0.80),
               (0.15, 0.95), (0.10, 1.10), (0.04, 1.18), (0.06, 1.22), (0.0, 1.26)],
    "queen": [(0.0, 0.0), (0.36, 0.0), (0.38, 0.07), (0.28, 0.17), (0.14, 0.38), (0.11, 0.85), (0.22, 0.92), (0.12, 0.98),
              (0.20, 1.20), (0.24, 1.30), (0.12, 1.32), (0.07, 1.40), (0.0, 1.44)],
    "king": [(0.0, 0.0), (0.37, 0.0), (0.39, 0.07), (0.29, 0.17), (0.15, 0.40), (0.12, 0.92), (0.23, 0.99), (0.13, 1.05),
             (0.19, 1.30), (0.22, 1.40), (0.08, 1.42), (0.05, 1.58), (0.0, 1.60)],
}


def _refine_profile(profile, rings):
    """Resamples a piecewise-linear profile to `rings` points with a little smoothing."""
    p = np.asarray(profile, np.float64)
    arc = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(p, axis=0), axis=1))])
    t = np.linspace(0, arc[-1], rings)
    out = np.stack([np.interp(t, arc, p[:, 0]), np.interp(t, arc, p[:, 1])], -1)
    k = np.array([0.25, 0.5, 0.25])
    sm = out.copy()
    for c in range(2):
        sm[1:-1, c] = np.convolve(out[:, c], k, mode="valid")
    return sm


def chess_scene(width: int = 1920, height: int = 1080, segments: int = 192, rings: int = 160, board_tess: int = 256,
                texture_size: int = 2048, seed: int = 0xAB600D) -> sc.SceneData:
    """BASELINE.json configs[1] stand-in ("ABeautifulGame-class"; the Khronos asset is not
    available offline, so every number here is declared by this builder, not by the asset).

    Defaults: 6 piece types turned on a lathe with `segments` x `rings` tessellation
    (2 * 192 * 159 = 61,056 triangles each), 32 piece instances, a 256 x 256-cell displaced board
    top (131,072 triangles), a frame of 4 boxes and a ground quad:
        32 * 61,056 + 131,072 + 48 + 2 = 2,084,914 instanced triangles, 15 metallic-roughness
    materials (rough dielectrics, glossy lacquer, brass / steel metals, two transmissive glass
    sets with volume attenuation), seven 2048^2 RGBA8 textures (colour sRGB, normal, ORM) from
    fixed-seed value noise, a directional sun plus 4 point lights and the constant sky."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size

    wood = value_noise(rs, n, 6, 4)
    grain = 0.5 + 0.5 * np.sin((np.arange(n)[None, :] / n * 40 + wood * 6) * 2 * np.pi)
    squares = ((np.indices((n, n)) // (n // 8)).sum(0) % 2).astype(F)
    light_sq = np.stack([0.80 + 0.1 * grain, 0.68 + 0.1 * grain, 0.45 + 0.1 * grain], -1)
    dark_sq = np.stack([0.22 + 0.08 * grain, 0.12 + 0.05 * grain, 0.07 + 0.03 * grain], -1)
    board_rgb = light_sq * squares[..., None] + dark_sq * (1 - squares[..., None])
    t_board = b.add_texture(rgba8(board_rgb[..., 0], board_rgb[..., 1], board_rgb[..., 2]), srgb=True)
    t_board_n = b.add_texture(normal_map_from_height(grain * 0.3 + wood, 6.0), srgb=False)
    t_board_orm = b.add_texture(rgba8(np.ones_like(wood), 0.25 + 0.35 * wood, np.zeros_like(wood)), srgb=False)
    marble = value_noise(rs, n, 7, 2)
    veins = np.abs(np.sin((marble * 5 + np.arange(n)[:, None] / n * 3) * np.pi)) ** 0.4
    t_white = b.add_texture(rgba8(0.92 * veins + 0.05, 0.90 * veins + 0.05, 0.84 * veins + 0.05), srgb=True)
    t_black = b.add_texture(rgba8(0.10 + 0.1 * (1 - veins), 0.10 + 0.09 * (1 - veins), 0.11 + 0.1 * (1 - veins)), srgb=True)
    t_piece_n = b.add_texture(normal_map_from_height(marble, 3.0), srgb=False)
    scratches = value_noise(rs, n, 6, 8)
    t_metal_orm = b.add_texture(rgba8(np.ones_like(scratches), 0.15 + 0.45 * scratches, 0.9 + 0.1 * scratches), srgb=False)

    mats = {
        "board": b.add_material_mr(roughness=0.9, color_idx=t_board, normal_idx=t_board_n, roughness_idx=t_board_orm),
        "frame": b.add_material_mr(color=(0.25, 0.14, 0.08, 1), roughness=0.45, normal_idx=t_board_n),
        "ground": b.add_material_mr(color=(0.55, 0.55, 0.58, 1), roughness=0.85),
        "white": b.add_material_mr(roughness=0.35, color_idx=t_white, normal_idx=t_piece_n),
        "white_gloss": b.add_material_mr(color=(0.95, 0.93, 0.88, 1), roughness=0.08, color_idx=t_white),
        "black": b.add_material_mr(roughness=0.30, color_idx=t_black, normal_idx=t_piece_n),
        "black_gloss": b.add_material_mr(color=(0.9, 0.9, 0.9, 1), roughness=0.06, color_idx=t_black),
        "brass": b.add_material_mr(color=(0.95, 0.76, 0.36, 1), roughness=0.6, metalness=1.0, roughness_idx=t_metal_orm,
                                   metallic_idx=t_metal_orm),
        "steel": b.add_material_mr(color=(0.77, 0.78, 0.80, 1), roughness=0.5, metalness=1.0, roughness_idx=t_metal_orm,
                                   metallic_idx=t_metal_orm),
        "gold": b.add_material_mr(color=(1.0, 0.80, 0.38, 1), roughness=0.18, metalness=1.0),
        "glass_clear": b.add_material_mr(color=(0.97, 0.98, 1.0, 1), roughness=0.02, transmission=1.0, ior=1.5),
        "glass_amber": b.add_material_mr(color=(1.0, 0.95, 0.85, 1), roughness=0.04, transmission=1.0, ior=1.52,
                                         attenuation_color=(0.95, 0.55, 0.15), attenuation_distance=0.6),
        "glass_smoke": b.add_material_mr(color=(0.9, 0.9, 0.95, 1), roughness=0.12, transmission=0.9, ior=1.45,
                                         attenuation_color=(0.35, 0.38, 0.45), attenuation_distance=0.5),
        "rubber": b.add_material_mr(color=(0.05, 0.05, 0.05, 1), roughness=0.95),
        "lacquer_red": b.add_material_mr(color=(0.65, 0.05, 0.04, 1), roughness=0.12),
    }
    assert len(mats) == 15

    geo = {k: b.add_geometry(*lathe(_refine_profile(p, rings), segments, uv_scale=(3.0, 2.0))) for k, p in _PIECES.items()}
    bump = value_noise(rs, 512, 5, 4).astype(np.float64)

    def board_height(x, z):
        ix = np.clip(((x / 8 + 0.5) * 511).astype(int), 0, 511)
        iz = np.clip(((z / 8 + 0.5) * 511).astype(int), 0, 511)
        return 0.004 * bump[iz, ix]

    g_board = b.add_geometry(*grid(board_tess, board_tess, 8, 8, uv_scale=1.0, height=board_height))
    g_box = b.add_geometry(*box(1, 1, 1))
    g_ground = b.add_geometry(*quad(60, 60, uv_scale=10))

    board = b.add_model([
        (g_board, mats["board"], None),
        (g_box, mats["frame"], translate(0, -0.15, 4.3) @ scale(9.2, 0.4, 0.6)),
        (g_box, mats["frame"], translate(0, -0.15, -4.3) @ scale(9.2, 0.4, 0.6)),
        (g_box, mats["frame"], translate(4.3, -0.15, 0) @ scale(0.6, 0.4, 8.0)),
        (g_box, mats["frame"], translate(-4.3, -0.15, 0) @ scale(0.6, 0.4, 8.0)),
    ])
    b.add_instance(board)
    b.add_instance(b.add_model([(g_ground, mats["ground"], None)]), translate(0, -0.36, 0))

    # one model per (piece type, material): the piece body plus nothing else, instanced per square
    def piece_model(kind, material):
        return b.add_model([(geo[kind], mats[material], None)])

    back = ["rook", "knight", "bishop", "queen", "king", "bishop", "knight", "rook"]
    white_special = {2: "glass_clear", 5: "glass_amber", 3: "gold", 0: "white_gloss", 7: "white_gloss"}
    black_special = {2: "glass_smoke", 5: "glass_smoke", 3: "steel", 4: "brass", 1: "lacquer_red", 6: "black_gloss"}
    models = {}

    def get(kind, material):
        if (kind, material) not in models:
            models[(kind, material)] = piece_model(kind, material)
        return models[(kind, material)]

    def square(file, rank):
        return (file - 3.5, 0.004, rank - 3.5)

    for f in range(8):
        jitter = rs.uniform(-0.06, 0.06, (4, 2))
        spin = rs.uniform(0, 360, 4)
        x, y, z = square(f, 0)
        b.add_instance(get(back[f], white_special.get(f, "white")), translate(x + jitter[0, 0], y, z + jitter[0, 1]) @ rotate_y(spin[0]))
        x, y, z = square(f, 1)
        b.add_instance(get("pawn", "white" if f % 3 else "white_gloss"), translate(x + jitter[1, 0], y, z + jitter[1, 1]) @ rotate_y(spin[1]))
        x, y, z = square(f, 6 if f != 4 else 4)  # one advanced pawn
        b.add_instance(get("pawn", "black" if f % 2 else "rubber"), translate(x + jitter[2, 0], y, z + jitter[2, 1]) @ rotate_y(spin[2]))
        x, y, z = square(f, 7)
        b.add_instance(get(back[f], black_special.get(f, "black")), translate(x + jitter[3, 0], y, z + jitter[3, 1]) @ rotate_y(spin[3]))

    b.set_directional_light((3.2, 3.0, 2.7), (-0.45, -1.0, -0.35))
    b.add_light((9.0, 7.0, 5.0), (-5.0, 3.0, 4.0), 1.0, 0.05, 0.08)
    b.add_light((4.0, 6.0, 9.0), (5.5, 2.5, -3.0), 1.0, 0.05, 0.10)
    b.add_light((6.0, 6.0, 6.0), (0.0, 5.0, 0.0), 1.0, 0.0, 0.12)
    b.add_light((5.0, 3.0, 2.0), (3.0, 1.2, 5.0), 1.0, 0.1, 0.2)
    eye = np.array([5.6, 4.2, -7.4])
    cam = camera_matrices(eye, np.array([0.2, 0.3, 0.2]) - eye, width, height, fov_deg=38)
    return b.build(cam, (width, height))


# ---------------------------------------------------------------------------------------------
# configs[2..4] stand-ins: every count below is declared by the builder
# ---------------------------------------------------------------------------------------------
def tube(center, radius, n_u: int, n_v: int, uv_scale=(1.0, 1.0)):
    """Closed tube (torus topology) around the closed curve center(t), t in [0, 2 pi), with
    radius(t, phi); seam vertices are duplicated so UVs run 0..uv_scale.  2 * n_u * n_v triangles."""
    t = np.linspace(0, 2 * np.pi, n_u + 1)
    phi = np.linspace(0, 2 * np.pi, n_v + 1)
    c = center(t)  # (n_u + 1, 3)
    e = 1e-4
    tangent = center(t + e) - center(t - e)
    tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
    # frame from the curve's distance to the y axis: continuous and periodic for the knots used here
    radial = c * np.array([1.0, 0.0, 1.0])
    radial /= np.maximum(np.linalg.norm(radial, axis=1, keepdims=True), 1e-9)
    side = np.cross(tangent, radial)
    side /= np.linalg.norm(side, axis=1, keepdims=True)
    up = np.cross(side, tangent)
    T, P = np.meshgrid(t, phi, indexing="ij")
    r = radius(T, P)

    def surface(rr):
        return c[:, None, :] + rr[..., None] * (np.cos(P)[..., None] * up[:, None, :] + np.sin(P)[..., None] * side[:, None, :])

    pos = surface(r)
    # normals from central differences of the parametric surface (radius varies along both directions)
    du = np.roll(pos[:-1], -1, 0) - np.roll(pos[:-1], 1, 0)
    du = np.concatenate([du, du[:1]], 0)
    dv = np.roll(pos[:, :-1], -1, 1) - np.roll(pos[:, :-1], 1, 1)
    dv = np.concatenate([dv, dv[:, :1]], 1)
    nrm = np.cross(dv, du)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-12)
    tan = du / np.maximum(np.linalg.norm(du, axis=-1, keepdims=True), 1e-12)
    bit = np.cross(nrm, tan)
    uv = np.stack([T / (2 * np.pi) * uv_scale[0], P / (2 * np.pi) * uv_scale[1]], -1)
    i = (np.arange(n_u)[:, None] * (n_v + 1) + np.arange(n_v)[None, :]).reshape(-1)
    idx = np.stack([i, i + n_v + 2, i + n_v + 1, i + n_v + 2, i, i + 1], -1).reshape(-1)  # CCW seen from outside
    flat = lambda a: a.reshape(-1, a.shape[-1])
    return _vertices(flat(pos), flat(uv), flat(nrm), flat(tan), flat(bit)), idx.astype(np.uint32)


def torus_knot(p: int, q: int, big: float, small: float):
    def center(t):
        r = big + small * np.cos(q * t)
        return np.stack([r * np.cos(p * t), small * np.sin(q * t), r * np.sin(p * t)], -1)

    return center


def dragon_scene(width: int = 1920, height: int = 1080, n_u: int = 4096, n_v: int = 64, cloth_tess: int = 256,
                 texture_size: int = 1024, seed: int = 0xD2A60) -> sc.SceneData:
    """BASELINE.json configs[2] stand-in ("DragonAttenuation-class"): one closed smooth transmissive
    body — a (2,3) torus-knot tube with a scaly radius modulation, 2 * n_u * n_v triangles
    (defaults: 524,288) — with Transmission 1, Ior 1.5, Roughness 0.02 and volume attenuation
    (AttenuationColor (0.92, 0.55, 0.15), AttenuationDistance 0.35 = the tube's thickness scale), on a
    displaced, textured cloth backdrop (2 * cloth_tess^2 triangles) and a floor quad; directional
    sun + 2 point lights, constant sky.  Stresses the BTDF lobe, Beer-Lambert attenuation
    (closestHit.rchit:123-128) and the TIR NaN restarts."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size
    weave = value_noise(rs, n, 6, 8)
    stripes = 0.5 + 0.5 * np.sin(np.arange(n)[None, :] / n * 2 * np.pi * 24)
    t_cloth = b.add_texture(rgba8(0.55 + 0.25 * stripes * weave, 0.12 + 0.1 * weave, 0.10 + 0.1 * weave), srgb=True)
    t_cloth_n = b.add_texture(normal_map_from_height(weave + 0.2 * stripes, 5.0), srgb=False)
    t_cloth_orm = b.add_texture(rgba8(np.ones_like(weave), 0.6 + 0.4 * weave, np.zeros_like(weave)), srgb=False)

    m_body = b.add_material_mr(color=(1.0, 1.0, 1.0, 1), roughness=0.02, transmission=1.0, ior=1.5,
                               attenuation_color=(0.92, 0.55, 0.15), attenuation_distance=0.35)
    m_cloth = b.add_material_mr(roughness=0.9, color_idx=t_cloth, normal_idx=t_cloth_n, roughness_idx=t_cloth_orm)
    m_floor = b.add_material_mr(color=(0.7, 0.7, 0.72, 1), roughness=0.6)

    scales = lambda T, P: 0.30 * (1.0 + 0.06 * np.cos(64 * T) * np.cos(8 * P) + 0.15 * np.sin(3 * T))
    g_body = b.add_geometry(*tube(torus_knot(2, 3, 1.6, 0.7), scales, n_u, n_v, uv_scale=(32.0, 2.0)))
    g_cloth = b.add_geometry(*grid(cloth_tess, cloth_tess, 14, 14, uv_scale=4.0,
                                   height=lambda x, z: 0.12 * np.sin(1.7 * x + 0.6 * np.sin(z)) * np.cos(1.1 * z)))
    g_floor = b.add_geometry(*quad(80, 80, uv_scale=16))
    b.add_instance(b.add_model([(g_body, m_body, None)]), translate(0, 1.45, 0) @ rotate_x(18))
    b.add_instance(b.add_model([(g_cloth, m_cloth, None)]), translate(0, 0.0, 0))
    b.add_instance(b.add_model([(g_cloth, m_cloth, None)]), translate(0, 5.0, 6.5) @ rotate_x(-78))
    b.add_instance(b.add_model([(g_floor, m_floor, None)]), translate(0, -0.2, 0))
    b.set_directional_light((4.0, 3.8, 3.4), (-0.35, -1.0, 0.45))
    b.add_light((12.0, 10.0, 8.0), (-4.0, 4.5, -3.0), 1.0, 0.05, 0.1)
    b.add_light((5.0, 7.0, 12.0), (4.5, 2.0, -4.0), 1.0, 0.05, 0.15)
    eye = np.array([0.5, 3.4, -7.2])
    cam = camera_matrices(eye, np.array([0.0, 1.3, 0.0]) - eye, width, height, fov_deg=40)
    return b.build(cam, (width, height))


def _leaf_texture(rs, n):
    """RGBA8 colour texture with (mostly) binary alpha from thresholded noise, coverage ~ 50 %; RGB is
    zeroed where alpha == 0 like the reference's importer does (TextureImporter.cpp:24-51)."""
    noise = value_noise(rs, n, 5, 8)
    alpha = np.clip((noise - np.median(noise)) * 24 + 0.5, 0, 1)
    tex = rgba8(0.10 + 0.25 * noise, 0.35 + 0.5 * noise, 0.08 + 0.1 * noise, alpha)
    tex[tex[..., 3] == 0, :3] = 0
    return tex


def atrium_scene(width: int = 3840, height: int = 2160, bays: int = 12, column_segments: int = 256, column_rings: int = 512,
                 floor_tess: int = 1024, cards_per_branch: int = 512, branches: int = 12288, texture_size: int = 1024,
                 seed: int = 0x5B0A2A) -> sc.SceneData:
    """BASELINE.json configs[3] stand-in ("Sponza-scale"): a two-aisle atrium of `bays` bays — fluted
    columns turned on a lathe (2 * column_segments * (column_rings - 1) triangles each, 4 rows),
    lintel boxes, a displaced floor and a vault (2 * floor_tess^2 triangles each) — all opaque, plus
    ALPHA-TESTED foliage: `branches` instances of a branch model holding `cards_per_branch` randomly
    oriented leaf cards (2 triangles each, IsOpaque = false, RGBA colour texture with ~50 % coverage)
    hanging as curtains between the columns.  Defaults:
        columns  4 * 12 * 2 * 256 * 511      = 12,558,336   (instanced, flattened by the core)
        floor + vault 2 * 2 * 1024^2         =  4,194,304
        lintels  4 * 12                      =         48
        foliage  12,288 * 512 * 2            = 12,582,912   (43 % of all triangles)
        total                                = 29,335,600 instanced triangles
    directional sun + constant sky + 6 point lights.  Stresses the any-hit stages (a4/a5)."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size
    stone = value_noise(rs, n, 7, 4)
    t_stone = b.add_texture(rgba8(0.62 + 0.25 * stone, 0.58 + 0.25 * stone, 0.50 + 0.25 * stone), srgb=True)
    t_stone_n = b.add_texture(normal_map_from_height(stone, 5.0), srgb=False)
    t_stone_orm = b.add_texture(rgba8(np.ones_like(stone), 0.55 + 0.4 * stone, np.zeros_like(stone)), srgb=False)
    t_leaf = b.add_texture(_leaf_texture(rs, n), srgb=True)
    tiles = ((np.indices((n, n)) // (n // 16)).sum(0) % 2).astype(F)
    t_floor = b.add_texture(rgba8(0.35 + 0.4 * tiles, 0.33 + 0.38 * tiles, 0.30 + 0.35 * tiles), srgb=True)

    m_stone = b.add_material_mr(roughness=0.9, color_idx=t_stone, normal_idx=t_stone_n, roughness_idx=t_stone_orm)
    m_floor = b.add_material_mr(roughness=0.35, color_idx=t_floor)
    m_leaf = b.add_material_mr(roughness=0.7, color_idx=t_leaf)
    m_curtain = b.add_material_mr(color=(0.9, 0.3, 0.25, 1), roughness=0.8, color_idx=t_leaf)

    # fluted column: radius modulated around the axis is not a lathe, so flutes go into the profile's
    # radius as fine rings (entasis + rings); base and capital included
    y = np.linspace(0, 1, column_rings)
    radius = 0.42 * (1 - 0.12 * y ** 2) * (1 + 0.015 * np.sin(y * 2 * np.pi * 40))
    radius[: column_rings // 24] *= 1.45
    radius[-column_rings // 24:] *= 1.5
    profile = np.stack([radius, y * 6.0], -1)
    g_column = b.add_geometry(*lathe(profile, column_segments, uv_scale=(3.0, 6.0)))
    g_box = b.add_geometry(*box(1, 1, 1))
    length = bays * 4.0
    g_floor = b.add_geometry(*grid(floor_tess, floor_tess, 16.0, length + 4, uv_scale=8.0,
                                   height=lambda x, z: 0.01 * np.sin(9 * x) * np.sin(9 * z)))
    g_vault = b.add_geometry(*grid(floor_tess, floor_tess, 16.0, length + 4, uv_scale=6.0,
                                   height=lambda x, z: -1.2 * np.cos(x * np.pi / 8) ** 2 - 0.05 * np.sin(3 * z)))
    # branch: cards scattered in a flat slab (a hanging curtain of leaves), random orientation
    cards_v, cards_i = [], []
    qv, qi = quad(0.22, 0.22)
    _tess = int(os.environ.get("PT_ATRIUM_CARD_TESS", "1"))  # experiment: tessellated cards (what reference splitting buys)
    if _tess > 1:
        qv, qi = grid(_tess, _tess, 0.22, 0.22)
    for k in range(cards_per_branch):
        m = (translate(*(rs.uniform(-0.5, 0.5, 3) * np.array([1.0, 1.0, 0.25]))) @ rotate_y(rs.uniform(0, 360)) @
             rotate_x(rs.uniform(40, 140)))
        v = qv.copy()
        p = np.concatenate([qv["position"], np.ones((len(qv), 1), F)], 1) @ m.T
        v["position"] = p[:, :3]
        for f in ("normal", "tangent", "bitangent"):
            v[f] = qv[f] @ m[:3, :3].T
        cards_v.append(v)
        cards_i.append(qi + len(qv) * k)
    g_branch = b.add_geometry(np.concatenate(cards_v), np.concatenate(cards_i), is_opaque=False)

    column = b.add_model([(g_column, m_stone, None)])
    lintel = b.add_model([(g_box, m_stone, None)])
    branch = [b.add_model([(g_branch, m_leaf, None)]), b.add_model([(g_branch, m_curtain, None)])]
    rows = (-6.0, -2.5, 2.5, 6.0)
    for bay in range(bays):
        z = (bay + 0.5) * 4.0 - length / 2
        for x in rows:
            b.add_instance(column, translate(x, 0, z))
    for x in rows:
        b.add_instance(lintel, translate(x, 6.25, 0) @ scale(1.2, 0.5, length))
    b.add_instance(b.add_model([(g_floor, m_floor, None)]))
    b.add_instance(b.add_model([(g_vault, m_stone, None)]), translate(0, 8.6, 0) @ rotate_x(180))
    # foliage curtains between neighbouring columns of every row, stacked vertically
    per_gap = max(1, branches // (len(rows) * bays))
    placed = 0
    for bay in range(bays):
        z = bay * 4.0 - length / 2 + 2.0
        for x in rows:
            for k in range(per_gap):
                if placed >= branches:
                    break
                s = rs.uniform(0.8, 1.2)
                b.add_instance(branch[placed & 1], translate(x + rs.uniform(-0.3, 0.3), rs.uniform(0.6, 5.8), z + 2.0 + rs.uniform(-1.4, 1.4)) @
                               rotate_y(rs.uniform(-20, 20) + 90) @ scale(s * 2.4, s * 1.6, s))
                placed += 1
    b.set_directional_light((5.0, 4.6, 4.0), (-0.5, -1.0, 0.25))
    for k in range(6):
        b.add_light((8.0, 6.5, 4.5), (0.0, 5.0, (k + 0.5) / 6 * length - length / 2), 1.0, 0.05, 0.12)
    eye = np.array([0.6, 2.0, -length / 2 + 1.0])
    cam = camera_matrices(eye, np.array([-0.4, 2.6, 0.0]) - eye, width, height, fov_deg=60)
    return b.build(cam, (width, height))


def street_scene(width: int = 3840, height: int = 2160, blocks: int = 24, facade_tess: int = 192, road_tess: int = 768,
                 lamps: int = 4096, lamp_segments: int = 32, texture_size: int = 1024, seed: int = 0xB157E0) -> sc.SceneData:
    """BASELINE.json configs[4] stand-in ("Bistro-scale"): a street of 2 * blocks displaced, textured
    facades (2 * facade_tess^2 triangles each), a displaced road, and `lamps` small emissive lamps
    (an emissive quad under a lathe-turned shade) strung along the street, plus the reference's
    maximum of 64 point lights and a dim directional light.  Reference semantics: emissive triangles
    are found by BSDF sampling only; NEE runs over the 65 analytic lights.  Defaults:
        facades 48 * 2 * 192^2 = 3,538,944; road 2 * 768^2 = 1,179,648;
        lamps 4096 * (2 + 2 * 32 * 15) = 3,940,352; total 8,658,944 instanced triangles."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size
    brick_noise = value_noise(rs, n, 6, 8)
    rows = (np.arange(n)[:, None] // (n // 32))
    bricks = (((np.arange(n)[None, :] + (rows % 2) * (n // 32)) // (n // 16)) + rows) % 3 / 2.0
    mortar = ((np.arange(n)[:, None] % (n // 32)) < 2) | (((np.arange(n)[None, :] + (rows % 2) * (n // 32)) % (n // 16)) < 2)
    shade_ = np.where(mortar, 0.75, 0.35 + 0.3 * bricks) * (0.8 + 0.2 * brick_noise)
    t_brick = b.add_texture(rgba8(shade_ * 1.1, shade_ * 0.62, shade_ * 0.5), srgb=True)
    t_brick_n = b.add_texture(normal_map_from_height(np.where(mortar, 0.0, 1.0) * 0.5 + brick_noise * 0.5, 5.0), srgb=False)
    windows = (((np.arange(n)[:, None] // (n // 8)) % 2 == 1) & ((np.arange(n)[None, :] // (n // 8)) % 2 == 1)).astype(F)
    glow = windows * (value_noise(rs, n, 3, 8) > 0.5)
    t_window_e = b.add_texture(rgba8(glow, 0.8 * glow, 0.45 * glow), srgb=True)
    asphalt = value_noise(rs, n, 7, 16)
    t_road = b.add_texture(rgba8(0.18 + 0.1 * asphalt, 0.18 + 0.1 * asphalt, 0.19 + 0.1 * asphalt), srgb=True)
    t_road_orm = b.add_texture(rgba8(np.ones_like(asphalt), 0.3 + 0.6 * asphalt, np.zeros_like(asphalt)), srgb=False)

    m_facade = b.add_material_mr(roughness=0.85, color_idx=t_brick, normal_idx=t_brick_n, emissive=(0, 0, 0),
                                 emissive_intensity=2.5, emissive_idx=t_window_e)
    m_road = b.add_material_mr(roughness=1.0, color_idx=t_road, roughness_idx=t_road_orm)
    m_shade = b.add_material_mr(color=(0.2, 0.2, 0.22, 1), roughness=0.4, metalness=1.0)
    lamp_colors = [(1.0, 0.85, 0.6), (1.0, 0.6, 0.3), (0.7, 0.85, 1.0), (1.0, 0.3, 0.3), (0.4, 1.0, 0.5)]
    m_bulbs = [b.add_material_mr(color=(1, 1, 1, 1), emissive=c, emissive_intensity=40.0) for c in lamp_colors]

    length = blocks * 6.0
    g_facade = b.add_geometry(*grid(facade_tess, facade_tess, 6.0, 9.0, uv_scale=2.0,
                                    height=lambda x, z: 0.04 * np.sin(7 * x) * np.sin(5 * z) + 0.1 * (np.abs(np.sin(2.1 * x)) > 0.93)))
    g_road = b.add_geometry(*grid(road_tess, road_tess, 12.0, length, uv_scale=12.0,
                                  height=lambda x, z: 0.015 * np.sin(5 * x) * np.sin(3 * z) - 0.03 * (x / 6) ** 2))
    g_bulb = b.add_geometry(*quad(0.12, 0.12))
    shade_profile = _refine_profile([(0.02, 0.0), (0.03, 0.05), (0.10, 0.08), (0.14, 0.16), (0.15, 0.18)], 16)
    g_shade = b.add_geometry(*lathe(shade_profile, lamp_segments))

    facade = b.add_model([(g_facade, m_facade, None)])
    b.add_instance(b.add_model([(g_road, m_road, None)]))
    for k in range(blocks):
        z = (k + 0.5) * 6.0 - length / 2
        b.add_instance(facade, translate(-6.0, 4.5, z) @ rotate_y(90) @ rotate_x(-90))   # facing +x
        b.add_instance(facade, translate(6.0, 4.5, z) @ rotate_y(-90) @ rotate_x(-90))   # facing -x
    lamp_models = [b.add_model([(g_bulb, m, translate(0, 0.17, 0) @ rotate_x(180)), (g_shade, m_shade, rotate_x(180) @ translate(0, -0.36, 0))])
                   for m in m_bulbs]
    for k in range(lamps):
        z = rs.uniform(-length / 2, length / 2)
        x = rs.uniform(-5.5, 5.5)
        ysag = 4.2 + 0.8 * (x / 5.5) ** 2 + rs.uniform(-0.1, 0.1)
        b.add_instance(lamp_models[k % len(lamp_models)], translate(x, ysag, z))
    for k in range(64):
        z = (k + 0.5) / 64 * length - length / 2
        b.add_light(tuple(4.0 * np.array(lamp_colors[k % 5])), ((-4.5, 4.5)[k & 1], 3.5, z), 1.0, 0.1, 0.25)
    b.set_directional_light((0.15, 0.17, 0.25), (-0.3, -1.0, 0.2))
    eye = np.array([1.2, 1.7, -length / 2 + 2.0])
    cam = camera_matrices(eye, np.array([-0.5, 2.4, 0.0]) - eye, width, height, fov_deg=62)
    return b.build(cam, (width, height))


def skinned_scene(width: int = 160, height: int = 120, segments: int = 24, rings: int = 40, bones: int = 4,
                  texture_size: int = 64, seed: int = 77) -> sc.SceneData:
    """A skinned column (a lathe-turned tube with `bones` bones along its axis, linear weights between
    neighbouring joints, up to 3 influences per vertex) bending over a static textured floor, next to a
    static sphere — the smallest scene that exercises skinning.comp's bone loop, the mixed static /
    animated buffers and the per-frame re-skin + rebuild.  Bone matrices come from bend_bones()."""
    rs = np.random.default_rng(seed)
    b = SceneBuilder()
    n = texture_size
    noise = value_noise(rs, n, 4)
    t_color = b.add_texture(rgba8(0.3 + 0.6 * noise, 0.5 + 0.3 * noise, 0.7 - 0.4 * noise), srgb=True)
    m_floor = b.add_material_mr(color=(0.8, 0.8, 0.8, 1), roughness=0.6, color_idx=t_color)
    m_skin = b.add_material_mr(color=(0.9, 0.5, 0.3, 1), roughness=0.35, metalness=0.2, color_idx=t_color)
    m_ball = b.add_material_mr(color=(0.9, 0.9, 1.0, 1), roughness=0.1, metalness=1.0)
    y = np.linspace(0, 2.0, rings)
    profile = np.stack([0.18 + 0.04 * np.sin(y * 9), y], -1)
    v, i = lathe(profile, segments, uv_scale=(2.0, 2.0))
    # weights: joint k sits at height k * 2 / bones; a vertex is bound to the joints around it (hat functions)
    h = v["position"][:, 1] / 2.0 * bones
    idx = np.zeros((len(v), 4), np.uint32)
    wgt = np.zeros((len(v), 4), F)
    k0 = np.clip(np.floor(h - 0.5).astype(int), 0, bones - 1)
    for slot, dk in enumerate((0, 1, 2)):
        k = np.clip(k0 + dk, 0, bones - 1)
        idx[:, slot] = k
        wgt[:, slot] = np.clip(1.0 - np.abs(h - 0.5 - k) , 0.0, 1.0)
    wgt[:, 0] += (wgt.sum(1) == 0)
    wgt /= wgt.sum(1, keepdims=True)
    g_tube = b.add_animated_geometry(v, i, idx, wgt)
    g_floor = b.add_geometry(*grid(8, 8, 6.0, 6.0, uv_scale=3.0))
    g_ball = b.add_geometry(*sphere(0.4, 16, 12))
    b.add_instance(b.add_model([(g_floor, m_floor, None)]))
    b.add_instance(b.add_model([(g_tube, m_skin, None)]), translate(-0.3, 0.0, 0.0))
    b.add_instance(b.add_model([(g_ball, m_ball, None)]), translate(0.9, 0.4, 0.3))
    b.set_directional_light((3.0, 2.8, 2.5), (-0.4, -1.0, -0.3))
    b.add_light((6.0, 5.0, 4.0), (1.5, 2.5, 1.5), 1.0, 0.1, 0.2)
    eye = np.array([0.4, 1.3, 3.2])
    cam = camera_matrices(eye, np.array([0.0, 0.9, 0.0]) - eye, width, height, fov_deg=50)
    s = b.build(cam, (width, height))
    s.bone_transforms = bend_bones(bones, 0.0)
    return s


def bend_bones(bones: int, angle_deg: float, length: float = 2.0) -> np.ndarray:
    """(bones, 12) bone matrices (3x4 row-major = glm::mat3x4 columns) of a chain along +y whose joint k rotates
    by angle_deg about z relative to its parent, joint k at height k * length / bones in the bind pose."""
    out = np.zeros((bones, 12), F)
    m = np.eye(4)
    for k in range(bones):
        pivot = translate(0.0, k * length / bones, 0.0).astype(np.float64)
        a = np.radians(angle_deg)
        rz = np.array([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        m = m @ pivot @ rz @ np.linalg.inv(pivot)
        out[k] = m[:3, :].astype(F).reshape(12)
    return out


WORKLOADS = {
    # name: (builder, BASELINE.json config index, default width, height, spp, depth)
    "chess": (chess_scene, 1, 1920, 1080, 256, 8),
    "dragon": (dragon_scene, 2, 1920, 1080, 1024, 16),
    "atrium": (atrium_scene, 3, 3840, 2160, 256, 8),
    "street": (street_scene, 4, 3840, 2160, 1024, 8),
}
