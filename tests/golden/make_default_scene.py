"""Generates tests/golden/default_scene.npz from the reference's own host code.

Runs oracle/_ref/scene_dump (the overlay build of the UNMODIFIED reference Scene / SceneManager /
ExampleScenes / TextureImporter / Camera sources, see oracle/ref_overlay/build.sh) for
"Test Scenes"/"Default" at 512x512 — BASELINE.json configs[0] — and stores the flattened PODs,
decoded textures and camera matrices as a compressed npz.  Needs /root/reference; the fixture it
writes is what travels.

    python tests/golden/make_default_scene.py
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "ref_overlay", "build.sh")])
    dump = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
    with tempfile.TemporaryDirectory() as tmp:
        raw = os.path.join(tmp, "default.ptscene")
        subprocess.check_call([dump, "Test Scenes", "Default", "512", "512", raw], cwd=tmp, stderr=subprocess.DEVNULL)
        scene = ptb200.SceneData.load_ptscene(raw)
    out = os.path.join(ROOT, "tests", "golden", "default_scene.npz")
    scene.save_npz(out)
    print(f"wrote {out}: {os.path.getsize(out)} bytes, {scene.instanced_triangle_count()} instanced triangles")


if __name__ == "__main__":
    main()
