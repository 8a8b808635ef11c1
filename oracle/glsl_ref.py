"""ctypes binding of oracle/_ref/libglsl_ref.so — the reference's own GLSL ray-tracing stages compiled
as C++ (oracle/ref_overlay/build_glsl.sh, glsl2cpp.py) — and of oracle/_ref/libglsl_comp_ref.so, its compute
stages (post-process chain, skinning) compiled the same way (glsl2cpp.py --compute).

TEST INFRASTRUCTURE, NOT PRODUCT: the pin of the CPU oracle (oracle/pt_oracle.cpp) to the reference's
shader text.  Only tests/ and tests/golden/make_glsl_vectors.py import this module.  The library can only
be BUILT where the reference checkout exists (this container); the built .so travels to the GPU box.
Traversal and texture filtering — which the reference leaves to the Vulkan implementation — are wired to
the oracle's (pto_trace_anyhit, pto_texture_sample, pto_sky_sample).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libglsl_ref.so")
_COMP_LIB_PATH = os.path.join(_HERE, "_ref", "libglsl_comp_ref.so")
REFERENCE_ROOT = "/root/reference"


class Callbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("trace", C.c_void_p), ("texture", C.c_void_p), ("sky", C.c_void_p), ("trace_flags", C.c_void_p)]


def build() -> str | None:
    """Builds the library when the reference checkout is present; returns its path or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "Path-Tracing", "Shaders")):
        subprocess.check_call([os.path.join(_HERE, "ref_overlay", "build_glsl.sh"), REFERENCE_ROOT], stdout=subprocess.DEVNULL)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available() -> bool:
    return build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("libglsl_ref.so is not built and the reference checkout is absent")
        L = C.CDLL(path)
        L.glr_scene_create.restype = C.c_void_p
        L.glr_scene_create.argtypes = [C.c_void_p, C.c_void_p]
        L.glr_scene_destroy.argtypes = [C.c_void_p]
        L.glr_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_void_p, C.c_int32]
        L.glr_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.glr_test_shading.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.glr_debug_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def test_shading(mode: int, inputs: np.ndarray) -> np.ndarray:
    inputs = np.ascontiguousarray(inputs, np.float32).reshape(-1, _oracle.TEST_IN[mode])
    out = np.zeros((inputs.shape[0], _oracle.TEST_OUT[mode]), np.float32)
    rc = lib().glr_test_shading(mode, inputs.ctypes.data, out.ctypes.data, inputs.shape[0])
    assert rc == 0, rc
    return out


class GlslScene:
    """The compiled GLSL stages bound to one scene; traversal / sampler served by `oracle_scene`."""

    def __init__(self, scene, oracle_scene):
        self._ora = oracle_scene  # keeps the callbacks' user pointer alive
        ol = _oracle.lib()
        cb = Callbacks(
            oracle_scene._h,
            C.cast(ol.pto_trace_anyhit, C.c_void_p).value,
            C.cast(ol.pto_texture_sample, C.c_void_p).value,
            C.cast(ol.pto_sky_sample, C.c_void_p).value,
            C.cast(ol.pto_trace_anyhit_flags, C.c_void_p).value,
        )
        desc, keep = scene.to_c()
        self._h = lib().glr_scene_create(C.addressof(desc), C.addressof(cb))
        del keep
        assert self._h, "glr_scene_create failed (animated geometry is not supported)"

    def close(self):
        if self._h:
            lib().glr_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, params, width, height, first_sample, frame_count, samples_per_frame=1, accum=None, threads=0):
        if accum is None:
            accum = np.zeros((height, width, 4), np.float32)
        p = params.to_c()
        rc = lib().glr_render(self._h, C.addressof(p), width, height, first_sample, frame_count, samples_per_frame,
                              accum.ctypes.data, threads)
        assert rc >= 0, rc
        return accum, rc

    def debug_render(self, params, width, height, mode=0, raygen_flags=0, hit_group_flags=0):
        """One frame of the compiled debug pipeline (Debug/debug*.r*): (H, W, 4) float32."""
        from importlib import import_module

        core = import_module("path-tracing_b200.core")
        d = core.DebugParams(core.DEBUG_MODES.index(mode) if isinstance(mode, str) else int(mode), raygen_flags, hit_group_flags)
        p = params.to_c()
        out = np.zeros((height, width, 4), np.float32)
        rc = lib().glr_debug_render(self._h, C.addressof(p), C.addressof(d), width, height, out.ctypes.data)
        assert rc == 0, rc
        return out

    def closest_hit(self, params, hits, rays6, payload_in):
        from importlib import import_module

        sc = import_module("path-tracing_b200.scene")
        hits = np.ascontiguousarray(hits, sc.HIT)
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        payload_in = np.ascontiguousarray(payload_in, np.float32).reshape(-1, 36)
        out = np.zeros_like(payload_in)
        p = params.to_c()
        rc = lib().glr_closest_hit(self._h, C.addressof(p), len(hits), hits.ctypes.data, rays6.ctypes.data,
                                   payload_in.ctypes.data, out.ctypes.data)
        assert rc == 0, rc
        return out


# ---- the COMPUTE stages (postprocess / bloom / composition / toneMapping / skinning .comp) --------------------------
_comp_lib = None


def comp_available() -> bool:
    build()
    return os.path.exists(_COMP_LIB_PATH)


def comp_lib():
    global _comp_lib
    if _comp_lib is None:
        if not comp_available():
            raise RuntimeError("libglsl_comp_ref.so is not built and the reference checkout is absent")
        L = C.CDLL(_COMP_LIB_PATH)
        L.glc_postprocess.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.glc_skin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        _comp_lib = L
    return _comp_lib


def postprocess(accum: np.ndarray, total_samples: int, exposure=1.0, bloom_threshold=1.0, bloom_intensity=0.1, tone_mapping_hdr=False):
    """The compiled compute stages over a (H, W, 4) float32 sum image: (bloom level 0 after the up-sampling chain,
    the post-process image after composition.comp, the same after toneMapping.comp), each (H, W, 4) float32 holding
    the RGBA16F values."""
    accum = np.ascontiguousarray(accum, np.float32)
    h, w = accum.shape[:2]
    outs = [np.zeros((h, w, 4), np.float32) for _ in range(3)]
    rc = comp_lib().glc_postprocess(accum.ctypes.data, w, h, int(total_samples), exposure, bloom_threshold, bloom_intensity,
                                    1 if tone_mapping_hdr else 0, *(o.ctypes.data for o in outs))
    assert rc == 0, rc
    return tuple(outs)


def skin_vertices(animated: np.ndarray, indices: np.ndarray, bone_transforms: np.ndarray) -> np.ndarray:
    """skinning.comp main() on animated[indices] -> (len(indices), 14) float32 vertices."""
    animated = np.ascontiguousarray(animated)
    assert animated.dtype.itemsize == 88, animated.dtype
    indices = np.ascontiguousarray(indices, np.uint32)
    bones = np.ascontiguousarray(bone_transforms, np.float32).reshape(-1, 12)
    out = np.zeros((len(indices), 14), np.float32)
    rc = comp_lib().glc_skin(animated.ctypes.data, len(animated), indices.ctypes.data, len(indices), bones.ctypes.data, len(bones),
                             out.ctypes.data)
    assert rc == 0, rc
    return out
