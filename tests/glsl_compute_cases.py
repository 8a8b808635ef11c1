"""Shared inputs of tests/test_oracle_vs_glsl_compute.py and tests/golden/make_glsl_compute_vectors.py: accumulation
images for the post-process chain and animated vertices / poses for skinning.comp."""
import importlib

import numpy as np

pkg = importlib.import_module("path-tracing_b200")
scenes = importlib.import_module("path-tracing_b200.scenes")

# (name, width, height, total samples, exposure, bloom threshold, bloom intensity)
POST_CASES = [
    ("defaults_96x54", 96, 54, 7, 1.0, 1.0, 0.1),
    ("strong_bloom_67x41", 67, 41, 3, 0.35, 0.6, 2.5),  # odd extents: the mip chain truncates, dispatches overhang
    ("tiny_5x3", 5, 3, 1, 1.0, 1.0, 0.1),  # fewer than 8 pixels: no bloom passes
    ("wide_256x32", 256, 32, 16, 2.0, 0.25, 1.0),
]
GOLDEN_POST = ("defaults_96x54", "strong_bloom_67x41", "tiny_5x3")
SKIN_ANGLES = (0.0, 12.0, 47.5)


def accumulation_image(w, h, total, seed):
    """A sum image with a wide dynamic range, exact zeros, and the NaN / Inf pixels postprocess.comp:24-27 marks."""
    rng = np.random.default_rng(seed)
    acc = (rng.random((h, w, 4), dtype=np.float32) ** 4 * np.float32(40.0 * total)).astype(np.float32)
    acc[..., 3] = total
    acc[h // 2, w // 3, 0] = np.nan
    acc[h // 3, w // 2, 1] = np.inf
    acc[0, 0, :3] = 0
    acc[h - 1, w - 1, :3] = 65504.0 * 4 * total  # beyond binary16: the RGBA16F store saturates to +Inf
    return acc


def post_case(name):
    for i, (n, w, h, total, exposure, threshold, intensity) in enumerate(POST_CASES):
        if n == name:
            return accumulation_image(w, h, total, 100 + i), total, exposure, threshold, intensity
    raise KeyError(name)


def skin_case(angle):
    """(animated vertices, indices in a scrambled order with repeats, bone transforms of the pose)."""
    s = scenes.skinned_scene(64, 48)
    bones = scenes.bend_bones(len(s.bone_transforms), angle)
    rng = np.random.default_rng(11)
    idx = rng.integers(0, len(s.animated_vertices), size=len(s.animated_vertices) + 37).astype(np.uint32)
    return s.animated_vertices, idx, np.ascontiguousarray(bones, np.float32)
