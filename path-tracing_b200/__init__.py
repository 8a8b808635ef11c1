"""B200-native path-tracing core — Python host side.

The directory name carries a hyphen (it is fixed by the project layout), so import it with
``importlib.import_module("path-tracing_b200")`` or through ``ptb200.py`` at the repo root.
"""
from .scene import *  # noqa: F401,F403
from .scene import SceneData, Texture, RenderParams  # noqa: F401
