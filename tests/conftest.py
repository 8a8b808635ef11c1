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
