#!/usr/bin/env python
"""Traversal-tail diagnostics of the bench workload: node-visit histogram per ray, warp loop
iterations in total / in drain mode.  usage: tools/tail_stats.py [spp]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
core = importlib.import_module("path-tracing_b200.core")
scenes = importlib.import_module("path-tracing_b200.scenes")
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H = 1920, 1080
scene = scenes.chess_scene(W, H)
params = scene.default_params(8)
r = core.Renderer(0)
r.update_scene_data(scene)
r.on_resize(W, H)
r.set_traversal_stats(True)
r.render(spp, params=params)
st = r.stats()
rays = st["rays_closest"] + st["rays_shadow"]
print("rays", rays, "ms", st["last_render_ms"], "iterations", st["wavefront_iterations"])
bins = ["<16", "<32", "<64", "<128", "<256", "<512", "<1024", ">=1024"]
hist = list(st["node_visit_hist"])
for b, h in zip(bins, hist):
    print(f"  visits {b:7s} {h:12d}  {100.0 * h / max(1, sum(hist)):7.3f}%")
print("warp iterations", st["warp_iterations"], "in drain mode", st["warp_drain_iterations"],
      f"({100.0 * st['warp_drain_iterations'] / max(1, st['warp_iterations']):.1f}%)", "max drain iterations of one warp",
      st["max_warp_drain_iterations"])
print("box tests / closest ray", st["box_tests_closest"] / max(1, st["rays_closest"]), "shadow", st["box_tests_shadow"] / max(1, st["rays_shadow"]))
