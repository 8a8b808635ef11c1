import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module("path-tracing_b200")
core = importlib.import_module("path-tracing_b200.core")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def default_scene():
    return pkg.SceneData.load_npz(os.path.join(ROOT, "tests", "golden", "default_scene.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def default_oracle(default_scene, oracle_mod):
    return oracle_mod.OracleScene(default_scene)


@pytest.fixture(scope="session")
def renderer():
    """The CUDA core through the C ABI.  Fails loudly (no fallback) if the library is missing."""
    r = core.Renderer(0)
    yield r
    r.close()


@pytest.fixture(scope="session")
def default_renderer(renderer, default_scene):
    renderer.update_scene_data(default_scene)
    return renderer


def diverged_path_fraction(renderer, oracle_scene, params, width, height, first_sample, count, tol=1e-4):
    """Fraction of (pixel, sample) pairs whose single-sample radiance differs between the CUDA core and the oracle by
    more than tol * max(1, |radiance|): paths that took a different route (another lobe, another side of an alpha
    or visibility test) because an intermediate value differed in its last bits."""
    bad = total = 0
    for s in range(first_sample, first_sample + count):
        renderer.on_resize(width, height)
        renderer.render(1, params=params, first_sample=s)
        a = renderer.read_accumulation()[..., :3]
        b, _ = oracle_scene.render(params, width, height, s, 1)
        b = b[..., :3]
        d = np.abs(a - b).max(-1)
        bad += int((d > tol * np.maximum(1.0, np.abs(b).max(-1))).sum())
        total += d.size
    return bad / total
