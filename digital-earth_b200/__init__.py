"""digital-earth_b200 -- B200-native drop-in for the Taichi kernels of AntonioFerreras/Digital-Earth.

Host side only: a ctypes binding of libde.so (include/de_api.h) behind the reference's own
`Renderer` surface (renderer.py), the 10-line `config.txt` scene format (earth_viewer.py:100-126,
213-236), screenshot output (earth_viewer.py:244-250) and the texture manifest (lib/textures.py).
All rendering arithmetic runs in hand-written sm_100a CUDA kernels; there is no CPU fallback.
"""
from .config import load_config, save_config  # noqa: F401
from .parameters import PathParameters, SceneParameters  # noqa: F401
from .renderer import Renderer  # noqa: F401
from .screenshot import save_screenshot, to_uint8_image  # noqa: F401
from .synth import make_textures  # noqa: F401
from . import textures  # noqa: F401

__all__ = ["Renderer", "load_config", "save_config", "save_screenshot", "to_uint8_image", "make_textures", "PathParameters",
           "SceneParameters", "textures"]
