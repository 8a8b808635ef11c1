"""Generates tests/golden/{box_textured,cylinder_engine}.npz with the reference's OWN importer.

oracle/ref_overlay/build_assimp.sh builds the assimp the reference vendors (glTF / OBJ importers only,
no CMake) and oracle/ref_overlay/build.sh then links the UNMODIFIED Path-Tracing/SceneImporter.cpp into
scene_dump, which loads the glTF files that ship inside vendor/assimp/test/models/glTF2, runs
SceneBuilder -> Scene::Update -> Camera and writes the flattened PODs.  Needs
/root/reference; the fixtures it writes are what travels.

    python tests/golden/make_imported_scenes.py
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

MODELS = "/root/reference/vendor/assimp/test/models/glTF2"
SCENES = {
    "box_textured": ("BoxTextured-glTF/BoxTextured.gltf", 256, 256),
    "cylinder_engine": ("2CylinderEngine-glTF-Binary/2CylinderEngine.glb", 320, 240),
}


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "ref_overlay", "build_assimp.sh")])
    subprocess.check_call([os.path.join(ROOT, "oracle", "ref_overlay", "build.sh")])
    dump = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
    for name, (rel, w, h) in SCENES.items():
        with tempfile.TemporaryDirectory() as tmp:
            raw = os.path.join(tmp, name + ".ptscene")
            subprocess.check_call([dump, "--file", os.path.join(MODELS, rel), str(w), str(h), raw], cwd=tmp,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            scene = ptb200.SceneData.load_ptscene(raw)
        if name == "cylinder_engine":
            # The file is modelled in millimetres (extent ~ 700 units) and its camera comes out of the reference's
            # importer at the origin, inside the engine block.  The reference's fixed epsilons (tmin = 1e-5, the
            # origin offsets of ray.glsl:93-131) assume metre-scale scenes: at 400 units one float ulp is 3e-5, and
            # whether a shadow ray escapes its own triangle is decided by the last bit.  So the one instance of the
            # model is placed in the world at 1:200, the way a user of the reference would place it, and the
            # camera (a render parameter, not scene data) frames its bounding box.
            import importlib

            import numpy as np

            scenes = importlib.import_module("path-tracing_b200.scenes")
            k = 1.0 / 200.0
            assert len(scene.instances) == 1
            scene.instances["transform"][0] = np.array([k, 0, 0, 0, 0, k, 0, 0, 0, 0, k, 0], np.float32)
            lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
            for rec in scene.mesh_records:
                g = scene.geometries[int(rec["geometry_index"])]
                m = np.asarray(scene.transforms[int(rec["transform_index"])], np.float64).reshape(3, 4)
                v = scene.vertices["position"][int(g["vertex_offset"]): int(g["vertex_offset"]) + int(g["vertex_length"])]
                pw = (v @ m[:, :3].T + m[:, 3]) * k
                lo, hi = np.minimum(lo, pw.min(0)), np.maximum(hi, pw.max(0))
            centre, radius = (lo + hi) / 2, np.linalg.norm(hi - lo) / 2
            eye = centre + radius * np.array([0.9, 0.6, 1.3])
            scene.view_inverse, scene.proj_inverse = scenes.camera_matrices(eye, centre - eye, w, h, fov_deg=45.0)
            scene.camera_extent = (w, h)
        out = os.path.join(ROOT, "tests", "golden", name + ".npz")
        scene.save_npz(out)
        print(f"wrote {out}: {os.path.getsize(out)} bytes, {scene.instanced_triangle_count()} instanced triangles, "
              f"{len(scene.textures)} textures, {len(scene.mr_materials)} MR materials")


if __name__ == "__main__":
    main()
