"""The C++ host shim under test: oracle/_ref/pt_headless is the reference's UNMODIFIED host code (SceneManager,
ExampleScenes, Scene, Camera, TextureImporter, compiled where it lies by oracle/ref_overlay/build.sh) driving the
core through path-tracing_b200/host/HeadlessRenderer — the binding a maintainer adds.  Its
accumulation image of `Test Scenes/Default` must equal the ctypes path's on the committed Default-scene fixture
(which the reference's own scene_dump wrote) BIT FOR BIT."""
import os
import subprocess

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
HEADLESS = os.path.join(conftest.ROOT, "oracle", "_ref", "pt_headless")


def run_headless(tmp_path, out_name, spp, env=None, size=128, bounces=8):
    out = tmp_path / out_name
    res = subprocess.run([HEADLESS, "Test Scenes", "Default", str(size), str(size), str(spp), str(bounces), str(out)],
                         capture_output=True, text=True, timeout=300, cwd=tmp_path, env=dict(os.environ, **(env or {})))
    assert res.returncode == 0, res.stderr[-2000:]
    return out, res.stdout


@pytest.fixture(scope="module")
def headless_available():
    if not os.path.exists(HEADLESS):
        pytest.skip("oracle/_ref/pt_headless is not built (needs the reference checkout: oracle/ref_overlay/build.sh)")


def test_shim_equals_ctypes_path_bit_for_bit(headless_available, tmp_path, default_renderer, default_scene):
    out, log = run_headless(tmp_path, "acc.f32", spp=2)
    shim = np.fromfile(out, np.float32).reshape(128, 128, 4)
    assert "2 spp" in log and "36 tris" in log
    r = default_renderer
    r.on_resize(128, 128)
    r.render(2, params=default_scene.default_params(bounce_count=8))
    assert np.array_equal(shim, r.read_accumulation())


def test_shim_samples_per_frame_and_output_formats(headless_available, tmp_path, default_renderer, default_scene):
    out, _ = run_headless(tmp_path, "acc4.f32", spp=4, env={"PT_SAMPLES_PER_FRAME": "2"})
    shim = np.fromfile(out, np.float32).reshape(128, 128, 4)
    r = default_renderer
    r.on_resize(128, 128)
    r.render_frames(2, 2, params=default_scene.default_params(bounce_count=8))
    assert np.array_equal(shim, r.read_accumulation())
    # OutputSaver's formats (Renderer/OutputSaver.h:16-19) through the shim: the files exist and carry the frame
    sizes = {}
    for ext in ("png", "jpg", "tga", "hdr"):
        path, _ = run_headless(tmp_path, "frame." + ext, spp=2)
        sizes[ext] = os.path.getsize(path)
    assert sizes["tga"] > 1000 and sizes["png"] > 1000 and sizes["jpg"] > 1000 and sizes["hdr"] > 128 * 128
    # the PNG of the shim is the PNG of the ctypes path (same pt_postprocess pixels)
    from PIL import Image

    r.on_resize(128, 128)
    r.render(2, params=default_scene.default_params(bounce_count=8))
    want = r.postprocess()
    assert np.array_equal(np.asarray(Image.open(tmp_path / "frame.png").convert("RGBA")), want)
    r.save_jpg(str(tmp_path / "mirror.jpg"))
    a = np.asarray(Image.open(tmp_path / "frame.jpg").convert("RGB"), np.int32)
    b = np.asarray(Image.open(tmp_path / "mirror.jpg").convert("RGB"), np.int32)
    # two JPEG encoders (stb in the shim as in the reference, Pillow in the mirror) of the same noisy 2-spp pixels: compared
    # after an 8 x 8 box filter (the block size), where quantisation noise averages out
    def blocks(x):
        return x.reshape(16, 8, 16, 8, 3).mean((1, 3))

    assert np.abs(blocks(a) - blocks(want[..., :3].astype(np.int32))).mean() < 2.0 and np.abs(blocks(a) - blocks(b)).mean() < 2.0
