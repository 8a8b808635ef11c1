"""pt_postprocess (CUDA) against the oracle's restatement of the reference chain, on the same accumulation image."""
import os

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
core = conftest.core


class _CudaArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3}


def _set_accumulation(r, image):
    """Writes a host image into the context's accumulation buffer (zero-copy torch view of pt_accum_device_ptr)."""
    import torch

    ptr, pitch, _ = r.accum_device_ptr()
    assert pitch == r.width * 16
    dev = torch.as_tensor(_CudaArray(ptr, (r.height, r.width, 4)), device="cuda:0")
    r.synchronize()
    dev.copy_(torch.from_numpy(np.ascontiguousarray(image, np.float32)))
    torch.cuda.synchronize()


def _compare(r, oracle_mod, acc, n, **kw):
    hdr = r.postprocess(hdr=True, total_samples=n, **kw)
    ref_hdr = oracle_mod.postprocess(acc, n, hdr=True, **kw)
    # every stage is the same sequence of unfused fp32 operations and half-precision stores
    # (bit for bit, except that the sign / payload of a NaN differs between x86 and the GPU)
    nan = np.isnan(ref_hdr)
    assert np.array_equal(np.isnan(hdr), nan)
    assert np.array_equal(hdr.view(np.uint32)[~nan], ref_hdr.view(np.uint32)[~nan]), np.abs(hdr - ref_hdr)[~nan].max()
    sdr = r.postprocess(hdr=False, total_samples=n, **kw)
    ref_sdr = oracle_mod.postprocess(acc, n, hdr=False, **kw)
    diff = np.abs(sdr.astype(np.int32) - ref_sdr.astype(np.int32))
    # exp / pow are evaluated in fp64 by two different math libraries: a last-bit difference can move a
    # value across a rounding boundary once in a long while
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-4, (diff.max(), (diff != 0).mean())
    assert (sdr[..., 3] == 255).all()
    return hdr, sdr


@pytest.mark.parametrize("size", [(256, 256), (250, 130), (67, 41)])
def test_rendered_image(default_renderer, default_scene, oracle_mod, size):
    """Default scene (emissive quad above the bloom threshold), power-of-two and odd frame sizes."""
    w, h = size
    p = default_scene.default_params(bounce_count=8)
    r = default_renderer
    r.on_resize(w, h)
    r.render(8, params=p)
    acc = r.read_accumulation()
    _compare(r, oracle_mod, acc, 8)
    _compare(r, oracle_mod, acc, 8, exposure=2.5, bloom_threshold=0.6, bloom_intensity=0.8)
    assert np.array_equal(r.read_accumulation(), acc), "pt_postprocess must not modify the accumulation buffer"


def test_synthetic_extremes(default_renderer, oracle_mod):
    """NaN / Inf markers, a single very bright pixel, and a value beyond the half range."""
    r = default_renderer
    w, h = 192, 128
    r.on_resize(w, h)
    rs = np.random.default_rng(11)
    acc = (rs.uniform(0, 1, (h, w, 4)) ** 4 * 6).astype(np.float32)
    acc[..., 3] = 1
    acc[5, 7, 0] = np.nan
    acc[50, 60, 1] = np.inf
    acc[64, 96, :3] = 30000.0
    _set_accumulation(r, acc)
    assert np.array_equal(r.read_accumulation(), acc, equal_nan=True)
    hdr, sdr = _compare(r, oracle_mod, acc, 3, bloom_intensity=0.3)
    assert np.isfinite(hdr).all()
    assert sdr[5, 7, 0] == 255 and sdr[50, 60, 1] == 255  # the markers survive (plus their neighbours' bloom)
    marker = r.postprocess(hdr=True, total_samples=3, bloom_intensity=0.0)
    assert tuple(marker[5, 7, :3]) == (5000.0, 0.0, 0.0) and tuple(marker[50, 60, :3]) == (0.0, 5000.0, 0.0)
    # A colour beyond 65504 becomes +inf in the RGBA16F images and the bloom chain spreads it (as NaN
    # once a zero weight or a zero intensity multiplies it) — the reference guards only its input.
    # The core reproduces that literally, Inf for Inf and NaN for NaN.
    acc[100, 20, 2] = 1e9
    _set_accumulation(r, acc)
    with np.errstate(invalid="ignore"):
        hdr, _ = _compare(r, oracle_mod, acc, 3, bloom_intensity=0.3)
        assert np.isinf(hdr[..., 2]).any() and np.isfinite(hdr[..., 0]).all()
        hdr, _ = _compare(r, oracle_mod, acc, 3, bloom_intensity=0.0)
        assert np.isnan(hdr[..., 2]).any() and np.isfinite(hdr[..., 0]).all()


def test_save_png_round_trip(default_renderer, default_scene, tmp_path):
    import zlib

    r = default_renderer
    r.on_resize(64, 48)
    r.render(2, params=default_scene.default_params(bounce_count=4))
    path = os.path.join(tmp_path, "out.png")
    r.save_png(path)
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[12:16] == b"IHDR"
    assert int.from_bytes(data[16:20], "big") == 64 and int.from_bytes(data[20:24], "big") == 48
    idat = data.index(b"IDAT")
    n = int.from_bytes(data[idat - 4:idat], "big")
    raw = np.frombuffer(zlib.decompress(data[idat + 4:idat + 4 + n]), np.uint8).reshape(48, 1 + 64 * 4)
    assert np.array_equal(raw[:, 1:].reshape(48, 64, 4), r.postprocess())


def test_errors(renderer):
    L = core.lib()
    p = core.PostProcessParams(1.0, 1.0, 0.1, 0)
    import ctypes as C

    buf = np.zeros(16, np.uint8)
    assert L.pt_postprocess(None, C.addressof(p), 1, 0, buf.ctypes.data, buf.nbytes) == -1
    r = core.Renderer(0)
    try:
        assert L.pt_postprocess(r._h, C.addressof(p), 1, 0, buf.ctypes.data, buf.nbytes) == -6  # PT_ERR_NO_TARGET
        r.on_resize(8, 8)
        assert L.pt_postprocess(r._h, C.addressof(p), 1, 0, buf.ctypes.data, buf.nbytes) == -1  # buffer too small
        assert L.pt_postprocess(r._h, C.addressof(p), 1, 7, buf.ctypes.data, 8 * 8 * 16) == -1  # unknown format
    finally:
        r.close()
