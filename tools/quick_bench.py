#!/usr/bin/env python
"""Quick A/B timing of the wavefront renderer on the bench workloads (no CPU legs, no parity block):
    [PT_CORE_LIB=scratch/variants/x/libpt_core.so] [PT_PLOC_RADIUS=32 ...] python tools/quick_bench.py \
        --workloads chess,atrium --spp 32 --reps 3 [--stats] [--tuning key=value,...] [--tag name]
Prints one JSON line per workload: Mrays/s of the best and the median repetition (library CUDA-event time of the
wavefront loop), build time, node count, and with --stats a separate counted render's boxes / triangles per ray."""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="chess")
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="per-kernel-class CUDA-event times of a separate ONE-pool render")
    ap.add_argument("--hash", action="store_true", help="sha1 of a 2-spp accumulation image (bit-identity of variants)")
    ap.add_argument("--tuning", default="")
    ap.add_argument("--tag", default=os.environ.get("PT_CORE_LIB", "in-tree"))
    a = ap.parse_args()
    scenes = importlib.import_module("path-tracing_b200.scenes")
    core = importlib.import_module("path-tracing_b200.core")
    for name in a.workloads.split(","):
        builder, _, w, h, _, depth = scenes.WORKLOADS[name]
        scene = builder(w, h)
        params = scene.default_params(depth)
        r = core.Renderer(0)
        for kv in filter(None, a.tuning.split(",")):
            k, v = kv.split("=")
            r.set_tuning(k, int(v))
        r.update_scene_data(scene)
        r.on_resize(w, h)
        r.render(max(1, a.spp // 4), params=params)  # warm-up
        rates = []
        for _ in range(a.reps):
            r.on_resize(w, h)
            r.render(a.spp, params=params)
            st = r.stats()
            rates.append((st["rays_closest"] + st["rays_shadow"]) / st["last_render_ms"] / 1e3)
        out = dict(tag=a.tag, workload=name, spp=a.spp, mrays_best=round(max(rates), 1), mrays_median=round(statistics.median(rates), 1),
                   build_ms=round(st["bvh_build_ms"], 1), nodes=st["bvh_node_count"], refs=st["bvh_reference_count"],
                   depth=st["bvh_max_depth"], iterations=st["wavefront_iterations"])
        if a.stats:
            r.set_traversal_stats(True)
            r.on_resize(w, h)
            r.render(max(1, a.spp // 8), params=params)
            st = r.stats()
            out.update(n_box_closest=round(st["box_tests_closest"] / max(1, st["rays_closest"]), 2),
                       n_tri_closest=round(st["tri_tests_closest"] / max(1, st["rays_closest"]), 2),
                       n_box_shadow=round(st["box_tests_shadow"] / max(1, st["rays_shadow"]), 2),
                       n_tri_shadow=round(st["tri_tests_shadow"] / max(1, st["rays_shadow"]), 2))
        if a.kernels:
            r.set_traversal_stats(False)
            r.set_tuning("pools", 1)
            r.set_kernel_timing(True)
            r.on_resize(w, h)
            r.render(a.spp, params=params)
            st = r.stats()
            out["one_pool_ms"] = {k: round(v, 2) for k, v in st["kernel_ms"].items()}
            r.set_kernel_timing(False)
            r.set_tuning("pools", 2)
        if a.hash:
            import hashlib
            r.set_traversal_stats(False)
            r.on_resize(w, h)
            r.render(2, params=params)
            out["sha1_2spp"] = hashlib.sha1(r.read_accumulation().tobytes()).hexdigest()[:16]
        print(json.dumps(out), flush=True)
        r.close()


if __name__ == "__main__":
    main()
