#!/usr/bin/env python
"""One render of a bench workload with ONE wavefront pool, for profiling under ncu:
    ncu --metrics ... python tools/render_once.py --workload chess --spp 8 --out gpurun_out/r.json
Writes the run's ray / launch counters so that tools/ncu_counters.py can turn per-kernel ncu totals into
per-ray figures.  A first untimed render warms the allocator up; the second is the profiled one (its kernel
launches are the LAST `launches` user kernels of the process)."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="chess")
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--pools", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    scenes = importlib.import_module("path-tracing_b200.scenes")
    core = importlib.import_module("path-tracing_b200.core")
    builder, _, w, h, _, depth = scenes.WORKLOADS[a.workload]
    scene = builder(w, h)
    params = scene.default_params(depth)
    r = core.Renderer(0)
    r.update_scene_data(scene)
    r.set_tuning("pools", a.pools)
    for _ in range(a.warm + 1):
        r.on_resize(w, h)
        r.render(a.spp, params=params)
    st = r.stats()
    out = {k: (int(v) if isinstance(v, (int,)) else v) for k, v in st.items() if k in (
        "rays_closest", "rays_shadow", "samples", "hits", "kernel_launches", "box_tests_closest", "tri_tests_closest",
        "box_tests_shadow", "tri_tests_shadow", "alpha_tests_closest", "alpha_tests_shadow", "texel_fetches", "wavefront_iterations")}
    out.update(workload=a.workload, spp=a.spp, width=w, height=h, depth=depth, renders=a.warm + 1)
    text = json.dumps(out)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")
    r.close()


if __name__ == "__main__":
    main()
