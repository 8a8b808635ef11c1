"""Reduced-tessellation builds of the BASELINE.json configs[1..4] stand-ins: the same builders,
materials, lights and cameras as the bench workloads, at sizes the CPU oracle renders in seconds."""
import importlib

scenes = importlib.import_module("path-tracing_b200.scenes")

W, H = 160, 120

SMALL = {
    # name: (builder call, bounce count of the config, expected instanced triangles)
    "chess": (lambda: scenes.chess_scene(W, H, segments=24, rings=20, board_tess=16, texture_size=128), 8,
              32 * 2 * 24 * 19 + 2 * 16 * 16 + 48 + 2),
    "dragon": (lambda: scenes.dragon_scene(W, H, n_u=512, n_v=24, cloth_tess=32, texture_size=128), 16,
               2 * 512 * 24 + 2 * 2 * 32 * 32 + 2),
    "atrium": (lambda: scenes.atrium_scene(W, H, bays=4, column_segments=24, column_rings=32, floor_tess=32,
                                           cards_per_branch=24, branches=96, texture_size=128), 8,
               4 * 4 * 2 * 24 * 31 + 4 * 12 + 2 * 2 * 32 * 32 + 96 * 24 * 2),
    "street": (lambda: scenes.street_scene(W, H, blocks=6, facade_tess=16, road_tess=32, lamps=128, lamp_segments=8,
                                           texture_size=128), 8,
               12 * 2 * 16 * 16 + 2 * 32 * 32 + 128 * (2 + 2 * 8 * 15)),
}
