"""ctypes binding of the CPU oracle (oracle/libpt_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")


class Counters(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "rays_closest",
            "rays_shadow",
            "samples",
            "hits",
            "box_tests_closest",
            "tri_tests_closest",
            "box_tests_shadow",
            "tri_tests_shadow",
            "alpha_tests_closest",
            "alpha_tests_shadow",
            "texel_fetches",
            "restarts",
        )
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compiles the oracle with the committed Makefile (gcc only, no CUDA)."""
    src = [os.path.join(_HERE, f) for f in ("pt_oracle.cpp", "pt_oracle_post.cpp", "pt_oracle.h", "glsl_math.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libpt_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.pto_scene_create.restype = C.c_void_p
        L.pto_scene_create.argtypes = [C.c_void_p]
        L.pto_scene_create_limits.restype = C.c_void_p
        L.pto_scene_create_limits.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.pto_scene_destroy.argtypes = [C.c_void_p]
        L.pto_scene_set_sampler.argtypes = [C.c_void_p, C.c_uint32]
        L.pto_scene_triangle_count.restype = C.c_uint64
        L.pto_scene_triangle_count.argtypes = [C.c_void_p]
        L.pto_render.argtypes = [
            C.c_void_p,
            C.c_void_p,
            C.c_uint32,
            C.c_uint32,
            C.c_uint32,
            C.c_uint32,
            C.c_void_p,
            C.c_uint32,
            C.c_void_p,
            C.c_int32,
            C.c_void_p,
        ]
        L.pto_render_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_uint32, C.c_void_p, C.c_int32, C.c_void_p]
        L.pto_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pto_first_hit_aov.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.pto_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.pto_trace_occlusion.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.pto_trace_closest_bruteforce_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.pto_test_shading.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pto_texture_info.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pto_texture_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.pto_texture_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32]
        L.pto_postprocess.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.pto_debug_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.pto_round_half.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.pto_skin_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def postprocess(accum: np.ndarray, total_samples: int, exposure=1.0, bloom_threshold=1.0, bloom_intensity=0.1,
                hdr: bool = False, tone_mapping_hdr: bool | None = None) -> np.ndarray:
    """The reference's post-process + output chain (pt_oracle_post.cpp) on a host (H, W, 4) float32 sum image:
    (H, W, 4) uint8 sRGB (png / jpg / tga outputs) or, with hdr, (H, W, 4) float32 (the .hdr output).
    tone_mapping_hdr overrides the tone-mapping mode the output format implies (float output of the SDR curve)."""
    from importlib import import_module

    core = import_module("path-tracing_b200.core")
    accum = np.ascontiguousarray(accum, np.float32)
    h, w = accum.shape[:2]
    tone_hdr = hdr if tone_mapping_hdr is None else tone_mapping_hdr
    p = core.PostProcessParams(exposure, bloom_threshold, bloom_intensity, 1 if tone_hdr else 0)
    out = np.zeros((h, w, 4), np.float32 if hdr else np.uint8)
    rc = lib().pto_postprocess(accum.ctypes.data, w, h, C.addressof(p), int(total_samples), 1 if hdr else 0, out.ctypes.data)
    assert rc == 0, rc
    return out


def skin_vertices(animated: np.ndarray, indices: np.ndarray, bone_transforms: np.ndarray) -> np.ndarray:
    """skinning.comp on animated[indices] -> (len(indices), 14) float32 vertices (pto_skin_vertices)."""
    animated = np.ascontiguousarray(animated)
    assert animated.dtype.itemsize == 88, animated.dtype
    indices = np.ascontiguousarray(indices, np.uint32)
    bones = np.ascontiguousarray(bone_transforms, np.float32).reshape(-1, 12)
    out = np.zeros((len(indices), 14), np.float32)
    rc = lib().pto_skin_vertices(animated.ctypes.data, indices.ctypes.data, len(indices), bones.ctypes.data, len(bones), out.ctypes.data)
    assert rc == 0, rc
    return out


def round_half(values: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(values, np.float32)
    out = np.empty_like(a)
    rc = lib().pto_round_half(a.ctypes.data, out.ctypes.data, a.size)
    assert rc == 0, rc
    return out


# record strides of pt(o)_test_shading, include/pt_core.h
TEST_IN = [4, 4, 4, 2, 1, 10, 11, 6, 3, 23, 21, 4, 42, 6, 2, 3, 30, 18, 12, 34, 35, 25, 22, 36, 3, 3]
TEST_OUT = [1, 1, 1, 1, 1, 4, 4, 3, 4, 4, 8, 9, 18, 3, 2, 9, 12, 6, 4, 12, 12, 3, 9, 12, 3, 3]


def test_shading(mode: int, inputs: np.ndarray) -> np.ndarray:
    inputs = np.ascontiguousarray(inputs, np.float32).reshape(-1, TEST_IN[mode])
    out = np.zeros((inputs.shape[0], TEST_OUT[mode]), np.float32)
    rc = lib().pto_test_shading(mode, inputs.ctypes.data, out.ctypes.data, inputs.shape[0])
    assert rc == 0, rc
    return out


class OracleScene:
    def __init__(self, scene, max_texture_size: int = 4096, texture_budget_mb: int = 0):
        from importlib import import_module

        self._sc = import_module("path-tracing_b200.scene")
        desc, keep = scene.to_c()
        self._h = lib().pto_scene_create_limits(C.addressof(desc), int(max_texture_size), int(texture_budget_mb) << 20)
        del keep
        assert self._h, "pto_scene_create failed"

    def close(self):
        if self._h:
            lib().pto_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sampler(self, max_anisotropy: int):
        """Maximum anisotropy of textureGrad: 1 = isotropic trilinear (default), 16 = the reference's sampler state."""
        rc = lib().pto_scene_set_sampler(self._h, int(max_anisotropy))
        assert rc == 0, rc

    @property
    def triangle_count(self) -> int:
        return int(lib().pto_scene_triangle_count(self._h))

    def render(self, params, width, height, first_sample, sample_count, accum=None, tiles=None, threads=0):
        if accum is None:
            accum = np.zeros((height, width, 4), np.float32)
        assert accum.dtype == np.float32 and accum.flags.c_contiguous
        p = params.to_c()
        cnt = Counters()
        tl = None if tiles is None else np.ascontiguousarray(tiles, self._sc.TILE)
        rc = lib().pto_render(
            self._h,
            C.addressof(p),
            width,
            height,
            first_sample,
            sample_count,
            None if tl is None else tl.ctypes.data,
            0 if tl is None else len(tl),
            accum.ctypes.data,
            threads,
            C.addressof(cnt),
        )
        assert rc == 0, rc
        return accum, cnt.as_dict()

    def render_frames(self, params, width, height, first_sample, frame_count, samples_per_frame, accum=None, tiles=None,
                      threads=0):
        """frame_count frames of samples_per_frame samples each (SampleCount > 1, Renderer.cpp:1688-1700)."""
        if accum is None:
            accum = np.zeros((height, width, 4), np.float32)
        p = params.to_c()
        cnt = Counters()
        tl = None if tiles is None else np.ascontiguousarray(tiles, self._sc.TILE)
        rc = lib().pto_render_frames(self._h, C.addressof(p), width, height, first_sample, frame_count, samples_per_frame,
                                     None if tl is None else tl.ctypes.data, 0 if tl is None else len(tl),
                                     accum.ctypes.data, threads, C.addressof(cnt))
        assert rc == 0, rc
        return accum, cnt.as_dict()

    def closest_hit(self, params, hits, rays6, payload_in):
        """closestHit.rchit on given hits; payloads as 36-word Shaders::Payload records."""
        hits = np.ascontiguousarray(hits, self._sc.HIT)
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        payload_in = np.ascontiguousarray(payload_in, np.float32).reshape(-1, 36)
        out = np.zeros_like(payload_in)
        p = params.to_c()
        rc = lib().pto_closest_hit(self._h, C.addressof(p), len(hits), hits.ctypes.data, rays6.ctypes.data,
                                   payload_in.ctypes.data, out.ctypes.data)
        assert rc == 0, rc
        return out

    def first_hit_aov(self, params, width, height):
        out = np.zeros(width * height, self._sc.HIT)
        p = params.to_c()
        rc = lib().pto_first_hit_aov(self._h, C.addressof(p), width, height, out.ctypes.data)
        assert rc == 0, rc
        return out.reshape(height, width)

    def debug_render(self, params, width, height, mode="color", raygen_flags=0, hit_group_flags=0):
        from importlib import import_module

        core = import_module("path-tracing_b200.core")
        d = core.DebugParams(core.DEBUG_MODES.index(mode) if isinstance(mode, str) else int(mode), raygen_flags, hit_group_flags)
        p = params.to_c()
        out = np.zeros((height, width, 4), np.float32)
        rc = lib().pto_debug_render(self._h, C.addressof(p), C.addressof(d), width, height, out.ctypes.data)
        assert rc == 0, rc
        return out

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, self._sc.RAY)
        out = np.zeros(len(rays), self._sc.HIT)
        rc = lib().pto_trace_closest(self._h, rays.ctypes.data, len(rays), out.ctypes.data)
        assert rc == 0, rc
        return out

    def trace_occlusion(self, rays):
        rays = np.ascontiguousarray(rays, self._sc.RAY)
        out = np.zeros(len(rays), np.uint8)
        rc = lib().pto_trace_occlusion(self._h, rays.ctypes.data, len(rays), out.ctypes.data)
        assert rc == 0, rc
        return out

    def trace_closest_bruteforce_f64(self, rays):
        rays = np.ascontiguousarray(rays, self._sc.RAY)
        out = np.zeros(len(rays), self._sc.HIT)
        t = np.zeros(len(rays), np.float64)
        rc = lib().pto_trace_closest_bruteforce_f64(self._h, rays.ctypes.data, len(rays), out.ctypes.data, t.ctypes.data)
        assert rc == 0, rc
        return out, t

    def texture_info(self, slot):
        w, h, l = C.c_uint32(), C.c_uint32(), C.c_uint32()
        rc = lib().pto_texture_info(self._h, slot, C.byref(w), C.byref(h), C.byref(l))
        assert rc == 0, rc
        return w.value, h.value, l.value

    def texture_level(self, slot, level):
        w, h, _ = self.texture_info(slot)
        w, h = max(1, w >> level), max(1, h >> level)
        out = np.zeros((h, w, 4), np.uint8)
        rc = lib().pto_texture_level(self._h, slot, level, out.ctypes.data)
        assert rc == 0, rc
        return out

    def texture_sample(self, slot, uv_ddx_ddy, use_grad=True):
        a = np.ascontiguousarray(uv_ddx_ddy, np.float32).reshape(-1, 6)
        out = np.zeros((a.shape[0], 4), np.float32)
        rc = lib().pto_texture_sample(self._h, slot, a.ctypes.data, out.ctypes.data, a.shape[0], 1 if use_grad else 0)
        assert rc == 0, rc
        return out
