"""Image comparison metrics used by the parity tests: re-exported from the package (path-tracing_b200/metrics.py,
where bench.py's parity block finds them too)."""
import importlib

_m = importlib.import_module("path-tracing_b200.metrics")
rel_mse, close_fraction, tonemap_srgb, flip, flip_map_ldr = _m.rel_mse, _m.close_fraction, _m.tonemap_srgb, _m.flip, _m.flip_map_ldr
