"""Host-side scene container: the POD arrays of ``pt_scene_desc`` (include/pt_core.h) as numpy
arrays, (de)serialisation, and the ctypes view handed to the C ABI.

The array layouts are byte-identical to the reference's host structs
(Path-Tracing/Shaders/ShaderTypes.incl:40-141, Path-Tracing/Scene.h:63-107); a scene produced by
the reference's own SceneBuilder (dumped by oracle/ref_overlay/scene_dump.cpp) and a procedural
one from scenes.py go through exactly the same path.
"""
from __future__ import annotations

import ctypes as C
import io
import struct
from dataclasses import dataclass, field

import numpy as np

# --- numpy dtypes mirroring the C structs -------------------------------------------------------

VERTEX = np.dtype(
    [("position", "<f4", 3), ("texcoords", "<f4", 2), ("normal", "<f4", 3), ("tangent", "<f4", 3), ("bitangent", "<f4", 3)]
)
ANIMATED_VERTEX = np.dtype(
    [("position", "<f4", 3), ("texcoords", "<f4", 2), ("normal", "<f4", 3), ("tangent", "<f4", 3), ("bitangent", "<f4", 3),
     ("bone_indices", "<u4", 4), ("bone_weights", "<f4", 4)]
)
GEOMETRY = np.dtype(
    [("vertex_offset", "<u4"), ("vertex_length", "<u4"), ("index_offset", "<u4"), ("index_length", "<u4"), ("is_opaque", "<u4")]
)
MESH_RECORD = np.dtype([("geometry_index", "<u4"), ("material_id", "<u4"), ("transform_index", "<u4")])
MODEL = np.dtype([("mesh_offset", "<u4"), ("mesh_count", "<u4")])
INSTANCE = np.dtype([("transform", "<f4", 12), ("model_index", "<u4")])
MATERIAL_MR = np.dtype(
    [
        ("emissive_color", "<f4", 3),
        ("emissive_intensity", "<f4"),
        ("color", "<f4", 4),
        ("roughness", "<f4"),
        ("metalness", "<f4"),
        ("ior", "<f4"),
        ("transmission", "<f4"),
        ("attenuation_color", "<f4", 3),
        ("attenuation_distance", "<f4"),
        ("pad", "<f4", 3),
        ("emissive_idx", "<u4"),
        ("color_idx", "<u4"),
        ("normal_idx", "<u4"),
        ("roughness_idx", "<u4"),
        ("metallic_idx", "<u4"),
    ]
)
MATERIAL_SG = np.dtype(
    [
        ("emissive_color", "<f4", 3),
        ("emissive_intensity", "<f4"),
        ("color", "<f4", 4),
        ("specular", "<f4", 3),
        ("glossiness", "<f4"),
        ("attenuation_color", "<f4", 3),
        ("attenuation_distance", "<f4"),
        ("ior", "<f4"),
        ("transmission", "<f4"),
        ("emissive_idx", "<u4"),
        ("color_idx", "<u4"),
        ("normal_idx", "<u4"),
        ("specular_idx", "<u4"),
        ("glossiness_idx", "<u4"),
        ("pad0", "<f4"),
    ]
)
POINT_LIGHT = np.dtype(
    [
        ("color", "<f4", 3),
        ("pad0", "<f4"),
        ("position", "<f4", 3),
        ("pad1", "<f4"),
        ("attenuation_constant", "<f4"),
        ("attenuation_linear", "<f4"),
        ("attenuation_quadratic", "<f4"),
        ("pad2", "<f4"),
    ]
)
DIRECTIONAL_LIGHT = np.dtype([("color", "<f4", 3), ("pad0", "<f4"), ("direction", "<f4", 3), ("pad1", "<f4")])
RAY = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("direction", "<f4", 3), ("tmax", "<f4")])
HIT = np.dtype([("instance", "<u4"), ("geometry", "<u4"), ("primitive", "<u4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
TILE = np.dtype([("x0", "<u4"), ("y0", "<u4"), ("x1", "<u4"), ("y1", "<u4")])

assert VERTEX.itemsize == 56 and MATERIAL_MR.itemsize == 96 and MATERIAL_SG.itemsize == 96
assert ANIMATED_VERTEX.itemsize == 88 and POINT_LIGHT.itemsize == 48 and DIRECTIONAL_LIGHT.itemsize == 32 and INSTANCE.itemsize == 52
assert HIT.itemsize == 24 and RAY.itemsize == 32

SCENE_TEXTURE_OFFSET = 9
NO_HIT = 0xFFFFFFFF
TEXTURE_RGBA8, TEXTURE_RGBAF32, TEXTURE_BC1, TEXTURE_BC3, TEXTURE_BC5 = 0, 1, 2, 3, 4
MATERIAL_TYPE_MR, MATERIAL_TYPE_SG, MATERIAL_TYPE_PHONG = 0, 1, 2
MISS_FLAGS_NONE, MISS_FLAGS_SKYBOX_2D, MISS_FLAGS_SKYBOX_CUBE = 0, 1, 2
HIT_FLAGS_NONE, HIT_FLAGS_DX_NORMAL_TEXTURES = 0, 1

# default texture slots, Path-Tracing/Shaders/ShaderTypes.incl:18-26
TEX_DEFAULT_COLOR, TEX_DEFAULT_NORMAL, TEX_DEFAULT_ROUGHNESS, TEX_DEFAULT_METALLIC, TEX_DEFAULT_EMISSIVE = 0, 1, 2, 3, 4
TEX_DEFAULT_SPECULAR, TEX_DEFAULT_GLOSSINESS, TEX_DEFAULT_SHININESS = 5, 6, 7


def material_id(index: int, material_type: int = MATERIAL_TYPE_MR) -> int:
    """CreateMaterialId, Path-Tracing/Shaders/ShaderTypes.incl:155-158."""
    return (index << 8) | material_type


# --- ctypes mirrors ------------------------------------------------------------------------------


class CTextureDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("srgb", C.c_uint32), ("pixels", C.c_void_p),
                ("levels", C.c_uint32)]


class CDirectionalLight(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("pad0", C.c_float), ("direction", C.c_float * 3), ("pad1", C.c_float)]


class CSceneDesc(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("vertex_count", C.c_uint64),
        ("indices", C.c_void_p),
        ("index_count", C.c_uint64),
        ("transforms", C.c_void_p),
        ("transform_count", C.c_uint32),
        ("geometries", C.c_void_p),
        ("geometry_count", C.c_uint32),
        ("mesh_records", C.c_void_p),
        ("mesh_record_count", C.c_uint32),
        ("models", C.c_void_p),
        ("model_count", C.c_uint32),
        ("instances", C.c_void_p),
        ("instance_count", C.c_uint32),
        ("mr_materials", C.c_void_p),
        ("mr_material_count", C.c_uint32),
        ("sg_materials", C.c_void_p),
        ("sg_material_count", C.c_uint32),
        ("phong_materials", C.c_void_p),
        ("phong_material_count", C.c_uint32),
        ("textures", C.c_void_p),
        ("texture_count", C.c_uint32),
        ("point_lights", C.c_void_p),
        ("point_light_count", C.c_uint32),
        ("directional_light", CDirectionalLight),
        ("skybox_2d", C.c_void_p),
        ("skybox_cube", C.c_void_p),
        ("geometry_is_animated", C.c_void_p),
        ("animated_vertices", C.c_void_p),
        ("animated_vertex_count", C.c_uint64),
        ("animated_indices", C.c_void_p),
        ("animated_index_count", C.c_uint64),
        ("bone_transforms", C.c_void_p),
        ("bone_count", C.c_uint32),
    ]


class CRenderParams(C.Structure):
    _fields_ = [
        ("view_inverse", C.c_float * 16),
        ("proj_inverse", C.c_float * 16),
        ("bounce_count", C.c_uint32),
        ("lens_radius", C.c_float),
        ("focal_distance", C.c_float),
        ("miss_flags", C.c_uint32),
        ("hit_flags", C.c_uint32),
    ]


@dataclass
class Texture:
    """One decoded level-0 image (pt_texture_desc), or — with bc_format set — the block-compressed mip
    chain of a .dds file: pixels is then the flat uint8 block data of `levels` levels, level 0 first."""

    pixels: np.ndarray  # (h, w, 4) uint8 or float32; BC: flat uint8 blocks
    srgb: bool = False
    bc_format: int = 0  # TEXTURE_BC1 / BC3 / BC5, 0 = uncompressed
    bc_extent: tuple = (0, 0)  # (width, height) of level 0 of a BC texture
    levels: int = 1

    @property
    def width(self) -> int:
        return int(self.bc_extent[0]) if self.bc_format else int(self.pixels.shape[1])

    @property
    def height(self) -> int:
        return int(self.bc_extent[1]) if self.bc_format else int(self.pixels.shape[0])

    @property
    def format(self) -> int:
        if self.bc_format:
            return self.bc_format
        return TEXTURE_RGBAF32 if self.pixels.dtype == np.float32 else TEXTURE_RGBA8


@dataclass
class RenderParams:
    """RaygenUniformData minus the sample counters (pt_render_params)."""

    view_inverse: np.ndarray  # 16 floats, column-major (glm)
    proj_inverse: np.ndarray
    bounce_count: int = 8
    lens_radius: float = 0.0
    focal_distance: float = 10.0
    miss_flags: int = MISS_FLAGS_NONE
    hit_flags: int = HIT_FLAGS_NONE

    def to_c(self) -> CRenderParams:
        p = CRenderParams()
        p.view_inverse[:] = [float(x) for x in np.asarray(self.view_inverse, np.float32).reshape(-1)]
        p.proj_inverse[:] = [float(x) for x in np.asarray(self.proj_inverse, np.float32).reshape(-1)]
        p.bounce_count = int(self.bounce_count)
        p.lens_radius = float(self.lens_radius)
        p.focal_distance = float(self.focal_distance)
        p.miss_flags = int(self.miss_flags)
        p.hit_flags = int(self.hit_flags)
        return p

    def nbytes(self) -> int:
        return C.sizeof(CRenderParams)


def _empty(dtype) -> np.ndarray:
    return np.zeros(0, dtype=dtype)


@dataclass
class SceneData:
    vertices: np.ndarray = field(default_factory=lambda: _empty(VERTEX))
    indices: np.ndarray = field(default_factory=lambda: _empty(np.uint32))
    transforms: np.ndarray = field(default_factory=lambda: np.array([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]], np.float32))
    geometries: np.ndarray = field(default_factory=lambda: _empty(GEOMETRY))
    mesh_records: np.ndarray = field(default_factory=lambda: _empty(MESH_RECORD))
    models: np.ndarray = field(default_factory=lambda: _empty(MODEL))
    instances: np.ndarray = field(default_factory=lambda: _empty(INSTANCE))
    mr_materials: np.ndarray = field(default_factory=lambda: _empty(MATERIAL_MR))
    sg_materials: np.ndarray = field(default_factory=lambda: _empty(MATERIAL_SG))
    phong_materials: np.ndarray = field(default_factory=lambda: _empty(MATERIAL_SG))
    textures: list = field(default_factory=list)  # list[Texture]; scene texture i -> slot 9 + i
    point_lights: np.ndarray = field(default_factory=lambda: _empty(POINT_LIGHT))
    directional_light: np.ndarray = field(default_factory=lambda: np.zeros((), DIRECTIONAL_LIGHT))
    skybox_2d: Texture | None = None
    skybox_cube: list | None = None  # six Textures: Front, Back, Up, Down, Left, Right (= +X -X +Y -Y +Z -Z)
    # skeletal animation: geometries flagged in geometry_is_animated address these buffers
    geometry_is_animated: np.ndarray | None = None  # (geometry_count,) uint32 or None
    animated_vertices: np.ndarray = field(default_factory=lambda: _empty(ANIMATED_VERTEX))
    animated_indices: np.ndarray = field(default_factory=lambda: _empty(np.uint32))
    bone_transforms: np.ndarray = field(default_factory=lambda: np.zeros((0, 12), np.float32))
    # default camera of the scene at `camera_extent` (not part of pt_scene_desc)
    camera_extent: tuple = (0, 0)
    view_inverse: np.ndarray | None = None
    proj_inverse: np.ndarray | None = None
    miss_flags: int = MISS_FLAGS_NONE
    hit_flags: int = HIT_FLAGS_NONE

    # -- sizes ------------------------------------------------------------------------------
    def instanced_triangle_count(self) -> int:
        total = 0
        for inst in self.instances:
            m = self.models[int(inst["model_index"])]
            for r in self.mesh_records[int(m["mesh_offset"]) : int(m["mesh_offset"]) + int(m["mesh_count"])]:
                total += int(self.geometries[int(r["geometry_index"])]["index_length"]) // 3
        return total

    def nbytes(self) -> int:
        n = sum(
            a.nbytes
            for a in (
                self.vertices,
                self.indices,
                self.transforms,
                self.geometries,
                self.mesh_records,
                self.models,
                self.instances,
                self.mr_materials,
                self.sg_materials,
                self.phong_materials,
                self.point_lights,
            )
        )
        n += sum(t.pixels.nbytes for t in self.textures)
        return n

    def default_params(self, bounce_count: int = 8) -> RenderParams:
        assert self.view_inverse is not None and self.proj_inverse is not None
        return RenderParams(
            self.view_inverse, self.proj_inverse, bounce_count=bounce_count, miss_flags=self.miss_flags, hit_flags=self.hit_flags
        )

    # -- C view -----------------------------------------------------------------------------
    def to_c(self):
        """Returns (CSceneDesc, keepalive) — keepalive must outlive the call using the desc."""
        keep = []

        def arr(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data if a.size else None, a.shape[0] if a.ndim else 1

        d = CSceneDesc()
        d.vertices, d.vertex_count = arr(self.vertices, VERTEX)
        d.indices, d.index_count = arr(self.indices, np.uint32)
        t = np.ascontiguousarray(self.transforms, np.float32).reshape(-1, 12)
        d.transforms, d.transform_count = arr(t, np.float32)
        d.geometries, d.geometry_count = arr(self.geometries, GEOMETRY)
        d.mesh_records, d.mesh_record_count = arr(self.mesh_records, MESH_RECORD)
        d.models, d.model_count = arr(self.models, MODEL)
        d.instances, d.instance_count = arr(self.instances, INSTANCE)
        d.mr_materials, d.mr_material_count = arr(self.mr_materials, MATERIAL_MR)
        d.sg_materials, d.sg_material_count = arr(self.sg_materials, MATERIAL_SG)
        d.phong_materials, d.phong_material_count = arr(self.phong_materials, MATERIAL_SG)
        d.point_lights, d.point_light_count = arr(self.point_lights, POINT_LIGHT)
        if self.geometry_is_animated is not None and np.any(self.geometry_is_animated):
            assert len(self.geometry_is_animated) == len(self.geometries)
            d.geometry_is_animated, _ = arr(self.geometry_is_animated, np.uint32)
            d.animated_vertices, d.animated_vertex_count = arr(self.animated_vertices, ANIMATED_VERTEX)
            d.animated_indices, d.animated_index_count = arr(self.animated_indices, np.uint32)
            bt = np.ascontiguousarray(self.bone_transforms, np.float32).reshape(-1, 12)
            d.bone_transforms, d.bone_count = arr(bt, np.float32)
        dl = np.ascontiguousarray(self.directional_light, DIRECTIONAL_LIGHT).reshape(())
        C.memmove(C.byref(d.directional_light), dl.tobytes(), 32)

        def tex_desc(tex: Texture) -> CTextureDesc:
            px = np.ascontiguousarray(tex.pixels)
            keep.append(px)
            return CTextureDesc(tex.width, tex.height, tex.format, 1 if tex.srgb else 0, px.ctypes.data, tex.levels)

        if self.textures:
            descs = (CTextureDesc * len(self.textures))(*[tex_desc(t) for t in self.textures])
            keep.append(descs)
            d.textures = C.cast(descs, C.c_void_p)
        d.texture_count = len(self.textures)
        if self.skybox_2d is not None:
            sky = tex_desc(self.skybox_2d)
            keep.append(sky)
            d.skybox_2d = C.cast(C.pointer(sky), C.c_void_p)
        if self.skybox_cube is not None:
            assert len(self.skybox_cube) == 6
            faces = (CTextureDesc * 6)(*[tex_desc(t) for t in self.skybox_cube])
            keep.append(faces)
            d.skybox_cube = C.cast(faces, C.c_void_p)
        return d, keep

    # -- serialisation ----------------------------------------------------------------------
    def save_npz(self, path) -> None:
        items = {
            "vertices": self.vertices,
            "indices": self.indices,
            "transforms": np.asarray(self.transforms, np.float32).reshape(-1, 12),
            "geometries": self.geometries,
            "mesh_records": self.mesh_records,
            "models": self.models,
            "instances": self.instances,
            "mr_materials": self.mr_materials,
            "sg_materials": self.sg_materials,
            "phong_materials": self.phong_materials,
            "point_lights": self.point_lights,
            "directional_light": self.directional_light,
            "camera_extent": np.asarray(self.camera_extent, np.uint32),
            "flags": np.asarray([self.miss_flags, self.hit_flags], np.uint32),
            "texture_srgb": np.asarray([t.srgb for t in self.textures], np.uint8),
        }
        if self.geometry_is_animated is not None:
            items["geometry_is_animated"] = np.asarray(self.geometry_is_animated, np.uint32)
            items["animated_vertices"] = self.animated_vertices
            items["animated_indices"] = self.animated_indices
            items["bone_transforms"] = np.asarray(self.bone_transforms, np.float32).reshape(-1, 12)
        if self.view_inverse is not None:
            items["view_inverse"] = np.asarray(self.view_inverse, np.float32)
            items["proj_inverse"] = np.asarray(self.proj_inverse, np.float32)
        for i, t in enumerate(self.textures):
            items[f"texture_{i}"] = t.pixels
        if self.skybox_2d is not None:
            items["skybox_2d"] = self.skybox_2d.pixels
            items["skybox_2d_srgb"] = np.asarray([self.skybox_2d.srgb], np.uint8)
        if self.skybox_cube is not None:
            items["skybox_cube"] = np.stack([t.pixels for t in self.skybox_cube])
            items["skybox_cube_srgb"] = np.asarray([self.skybox_cube[0].srgb], np.uint8)
        np.savez_compressed(path, **items)

    @staticmethod
    def load_npz(path) -> "SceneData":
        z = np.load(path)
        s = SceneData()
        for k in (
            "vertices",
            "indices",
            "transforms",
            "geometries",
            "mesh_records",
            "models",
            "instances",
            "mr_materials",
            "sg_materials",
            "phong_materials",
            "point_lights",
            "directional_light",
        ):
            setattr(s, k, z[k])
        if "geometry_is_animated" in z:
            for k in ("geometry_is_animated", "animated_vertices", "animated_indices", "bone_transforms"):
                setattr(s, k, z[k])
        s.camera_extent = tuple(int(x) for x in z["camera_extent"])
        s.miss_flags, s.hit_flags = (int(x) for x in z["flags"])
        if "view_inverse" in z:
            s.view_inverse = z["view_inverse"]
            s.proj_inverse = z["proj_inverse"]
        srgb = z["texture_srgb"]
        s.textures = [Texture(z[f"texture_{i}"], bool(srgb[i])) for i in range(len(srgb))]
        if "skybox_2d" in z:
            s.skybox_2d = Texture(z["skybox_2d"], bool(z["skybox_2d_srgb"][0]))
        if "skybox_cube" in z:
            s.skybox_cube = [Texture(f, bool(z["skybox_cube_srgb"][0])) for f in z["skybox_cube"]]
        return s

    @staticmethod
    def load_ptscene(path) -> "SceneData":
        """Reads the chunked dump written by oracle/ref_overlay/scene_dump.cpp."""
        data = open(path, "rb").read()
        assert data[:8] == b"PTSCENE1", "not a PTSCENE1 file"
        f = io.BytesIO(data[8:])
        s = SceneData()
        pending = None
        dtypes = {
            "vertices": VERTEX,
            "indices": np.uint32,
            "geometries": GEOMETRY,
            "mesh_records": MESH_RECORD,
            "models": MODEL,
            "instances": INSTANCE,
            "mr_materials": MATERIAL_MR,
            "sg_materials": MATERIAL_SG,
            "phong_materials": MATERIAL_SG,
            "point_lights": POINT_LIGHT,
            "geometry_is_animated": np.uint32,
            "animated_vertices": ANIMATED_VERTEX,
            "animated_indices": np.uint32,
        }
        while True:
            tag = f.read(24)
            if len(tag) < 24:
                break
            name = tag.split(b"\0", 1)[0].decode()
            (nbytes,) = struct.unpack("<Q", f.read(8))
            payload = f.read(nbytes)
            if name in dtypes:
                setattr(s, name, np.frombuffer(payload, dtype=dtypes[name]).copy())
            elif name == "transforms":
                s.transforms = np.frombuffer(payload, np.float32).reshape(-1, 12).copy()
            elif name == "bone_transforms":
                s.bone_transforms = np.frombuffer(payload, np.float32).reshape(-1, 12).copy()
            elif name == "directional_light":
                s.directional_light = np.frombuffer(payload, DIRECTIONAL_LIGHT)[0].copy()
            elif name == "texture_info":
                pending = struct.unpack("<4I", payload)
            elif name == "texture_pixels":
                w, h, fmt, srgb = pending
                dt = np.float32 if fmt == TEXTURE_RGBAF32 else np.uint8
                s.textures.append(Texture(np.frombuffer(payload, dt).reshape(h, w, 4).copy(), bool(srgb)))
            elif name == "camera_extent":
                s.camera_extent = struct.unpack("<2I", payload)
            elif name == "view_inverse":
                s.view_inverse = np.frombuffer(payload, np.float32).copy()
            elif name == "proj_inverse":
                s.proj_inverse = np.frombuffer(payload, np.float32).copy()
            elif name == "flags":
                sky, dx = struct.unpack("<2I", payload)
                s.miss_flags = MISS_FLAGS_SKYBOX_2D if sky else MISS_FLAGS_NONE
                s.hit_flags = HIT_FLAGS_DX_NORMAL_TEXTURES if dx else HIT_FLAGS_NONE
        return s
