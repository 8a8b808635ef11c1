#!/usr/bin/env python3
"""Generates tests/golden/glsl_compute_vectors.npz: outputs of the REFERENCE'S OWN compute shaders —
postprocess.comp, bloomDownsample.comp, bloomUpsample.comp, composition.comp, toneMapping.comp, skinning.comp compiled
as C++ by oracle/ref_overlay/build_glsl.sh into oracle/_ref/libglsl_comp_ref.so — on the seeded inputs of
tests/glsl_compute_cases.py.  Needs the reference checkout (/root/reference): run in the build container; the vectors
then travel with the repository so that the oracle stays pinned on machines without the reference.

  post_<case>_{bloom0, composed, final}    RGBA16F values as float32
  skin_<angle>                             skinned vertices (14 floats each)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import glsl_compute_cases as cc  # noqa: E402
from oracle import glsl_ref  # noqa: E402


def main():
    assert glsl_ref.comp_available(), "reference checkout missing: cannot build libglsl_comp_ref.so"
    out = {}
    for name in cc.GOLDEN_POST:
        acc, total, exposure, threshold, intensity = cc.post_case(name)
        b0, composed, final = glsl_ref.postprocess(acc, total, exposure, threshold, intensity, tone_mapping_hdr=False)
        out[f"post_{name}_bloom0"], out[f"post_{name}_composed"], out[f"post_{name}_final"] = b0, composed, final
    for angle in cc.SKIN_ANGLES:
        out[f"skin_{angle}"] = glsl_ref.skin_vertices(*cc.skin_case(angle))
    path = os.path.join(HERE, "glsl_compute_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
