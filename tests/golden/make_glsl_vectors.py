#!/usr/bin/env python3
"""Generates tests/golden/glsl_vectors.npz: outputs of the REFERENCE'S OWN shader code — the GLSL stages
compiled as C++ by oracle/ref_overlay/build_glsl.sh into oracle/_ref/libglsl_ref.so — on seeded inputs.

Needs the reference checkout (/root/reference): run in the build container; the vectors then travel with
the repository so that the oracle stays pinned on machines without the reference (the GPU box).

  per function (every PT_TEST_* mode):  in_<mode>, out_<mode>       256 records each
  closestHit.rchit payloads:            chit_<scene>_{hits,rays,in,out}
  whole pipeline (raygen.rgen main):    img_<case>                  accumulation images
  debug pipeline (debugRaygen main):    dbg_<scene>_<mode>_<raygen flags>_<hit-group flags>
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import glsl_cases as gc  # noqa: E402
import unit_inputs as ui  # noqa: E402
from oracle import glsl_ref, oracle  # noqa: E402


def main():
    oracle.build()
    out = {}
    for mode in range(len(oracle.TEST_IN)):
        x = ui.inputs(mode, gc.GOLDEN_RECORDS, seed=7)
        out[f"in_{mode}"] = x
        out[f"out_{mode}"] = glsl_ref.test_shading(mode, x)
    default_scene = gc.pkg.SceneData.load_npz(os.path.join(HERE, "default_scene.npz"))
    for name, (scene, params, w, h, first, frames, spp) in gc.stage_scenes(default_scene).items():
        ora = oracle.OracleScene(scene)
        g = glsl_ref.GlslScene(scene, ora)
        img, spinning = g.render(params, w, h, first, frames, spp)
        assert spinning == 0
        out[f"img_{name}"] = img
        if name in ("default", "feature"):
            hits, rays, pin = gc.closest_hit_inputs(ora, oracle, params, w, h)
            out[f"chit_{name}_hits"] = hits
            out[f"chit_{name}_rays"] = rays
            out[f"chit_{name}_in"] = pin
            out[f"chit_{name}_out"] = g.closest_hit(params, hits, rays, pin)
    cases = gc.stage_scenes(default_scene)
    for scene_name, mode, rf, hf in gc.DEBUG_GOLDEN:
        scene, params, w, h = cases[scene_name][:4]
        g = glsl_ref.GlslScene(scene, oracle.OracleScene(scene))
        out[gc.debug_key(scene_name, mode, rf, hf)] = g.debug_render(params, w, h, mode, rf, hf)
    path = os.path.join(HERE, "glsl_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
