"""Struct layouts of the data contract.  The reference pins std430 == C++ layout with literal
field values in Path-Tracing-Tests/PaddingTest.cpp:11-209; the same literals are pushed through
the numpy dtypes (path-tracing_b200/scene.py) that back the C ABI structs of include/pt_core.h."""
import ctypes as C
import struct

import numpy as np

import conftest

sc = conftest.pkg.scene


def test_sizes():
    # sizeof() values established by compiling the reference headers
    assert sc.VERTEX.itemsize == 56
    assert sc.MATERIAL_MR.itemsize == 96 and sc.MATERIAL_SG.itemsize == 96
    assert sc.POINT_LIGHT.itemsize == 48 and sc.DIRECTIONAL_LIGHT.itemsize == 32
    assert sc.GEOMETRY.itemsize == 20 and sc.MESH_RECORD.itemsize == 12 and sc.MODEL.itemsize == 8
    assert sc.INSTANCE.itemsize == 52 and sc.RAY.itemsize == 32 and sc.HIT.itemsize == 24
    assert C.sizeof(sc.CRenderParams) == 148  # 2 mat4 + bounce, lens, focal + the two specialisation constants
    assert C.sizeof(sc.CTextureDesc) == 32  # + levels (block-compressed mip chains), padded to the pointer alignment


def test_metallic_roughness_padding_literals():
    """PaddingTest.MetallicRoughnessMaterial, first instance (PaddingTest.cpp:16-21)."""
    m = np.zeros((), sc.MATERIAL_MR)
    m["emissive_color"] = (1.1, 2.2, 3.3)
    m["emissive_intensity"] = 4.4
    m["color"] = (1.0, 2.0, 3.0, 12.0)
    m["roughness"], m["metalness"], m["ior"], m["transmission"] = 4.0, 5.0, 1.5, 2.5
    m["attenuation_color"] = (3.5, 4.5, 5.5)
    m["attenuation_distance"] = 6.5
    m["emissive_idx"], m["color_idx"], m["normal_idx"], m["roughness_idx"], m["metallic_idx"] = 10, 1, 2, 3, 4
    raw = m.tobytes()
    f = struct.unpack("<16f", raw[:64])
    assert np.allclose(f[:4], [1.1, 2.2, 3.3, 4.4]) and f[4:8] == (1.0, 2.0, 3.0, 12.0)
    assert f[8:12] == (4.0, 5.0, 1.5, 2.5) and f[12:16] == (3.5, 4.5, 5.5, 6.5)
    assert struct.unpack("<5I", raw[76:96]) == (10, 1, 2, 3, 4)


def test_specular_glossiness_padding_literals():
    """PaddingTest.SpecularGlossinessMaterial, first instance (PaddingTest.cpp:66-71)."""
    m = np.zeros((), sc.MATERIAL_SG)
    m["emissive_color"], m["emissive_intensity"] = (1.1, 2.2, 3.3), 4.4
    m["color"] = (1.0, 2.0, 3.0, 12.0)
    m["specular"], m["glossiness"] = (4.0, 5.0, 6.0), 7.0
    m["attenuation_color"], m["attenuation_distance"] = (1.1, 1.2, 1.3), 1.4
    m["ior"], m["transmission"] = 1.5, 1.6
    m["emissive_idx"], m["color_idx"], m["normal_idx"], m["specular_idx"], m["glossiness_idx"] = 10, 1, 2, 3, 4
    raw = m.tobytes()
    f = struct.unpack("<18f", raw[:72])
    assert f[8:12] == (4.0, 5.0, 6.0, 7.0) and np.allclose(f[12:18], [1.1, 1.2, 1.3, 1.4, 1.5, 1.6])
    assert struct.unpack("<5I", raw[72:92]) == (10, 1, 2, 3, 4)


def test_light_layouts():
    p = np.zeros((), sc.POINT_LIGHT)
    p["color"], p["position"] = (1, 2, 3), (4, 5, 6)
    p["attenuation_constant"], p["attenuation_linear"], p["attenuation_quadratic"] = 7, 8, 9
    f = struct.unpack("<12f", p.tobytes())
    assert f[0:3] == (1, 2, 3) and f[4:7] == (4, 5, 6) and f[8:11] == (7, 8, 9)
    d = np.zeros((), sc.DIRECTIONAL_LIGHT)
    d["color"], d["direction"] = (1, 2, 3), (4, 5, 6)
    f = struct.unpack("<8f", d.tobytes())
    assert f[0:3] == (1, 2, 3) and f[4:7] == (4, 5, 6)


def test_material_id_packing():
    assert sc.material_id(7, sc.MATERIAL_TYPE_PHONG) == (7 << 8) | 2


def test_default_scene_matches_reference_probe(default_scene):
    """Counts printed by the reference's own SceneBuilder for Test Scenes/Default."""
    s = default_scene
    assert len(s.vertices) == 48 and len(s.indices) == 72 and len(s.geometries) == 12
    assert len(s.transforms) == 1 and len(s.mr_materials) == 10 and len(s.textures) == 4
    assert len(s.models) == 4 and [int(m["mesh_offset"]) for m in s.models] == [0, 5, 11, 17]
    assert len(s.mesh_records) == 18 and len(s.instances) == 4 and s.instanced_triangle_count() == 36
    assert len(s.point_lights) == 0
    assert tuple(s.directional_light["color"]) == (0, 0, 0) and tuple(s.directional_light["direction"]) == (0, -1, 0)
    assert [t.pixels.shape[:2] for t in s.textures] == [(1024, 1024), (512, 512), (512, 512), (2024, 2024)]
    assert np.allclose(s.instances[0]["transform"], [2, 0, 0, -4.5, 0, 2, 0, 1, 0, 0, 2, 0])
    assert np.allclose(s.view_inverse, [0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, 0, 3, 1, 0, 1], atol=1e-6)
    assert np.allclose(s.proj_inverse, [0.414214, 0, 0, 0, 0, 0.414214, 0, 0, 0, 0, 0, 9.99, 0, 0, 1, 0.01], atol=1e-5)
