"""mtgs_b200 -- B200-native (sm_100a) Gaussian-splat rasterizer behind MTGS's gsplat call surface.

Public API (mirrors the two upstream entry points MTGS imports):

    from mtgs_b200.rendering import rasterization            # gsplat.rendering.rasterization
    from mtgs_b200.cuda._wrapper import spherical_harmonics  # gsplat.cuda._wrapper.spherical_harmonics

``install_as_gsplat()`` registers these modules under the ``gsplat`` names so the unmodified MTGS imports
(mtgs/scene_model/mtgs_scene_graph.py:20-23, gaussian_model/vanilla_gaussian_splatting.py:15-18) resolve here.
"""
from __future__ import annotations

import sys
import types

__version__ = "0.1.0"


def install_as_gsplat(force: bool = False) -> None:
    """Make ``import gsplat.rendering`` / ``gsplat.cuda._wrapper`` resolve to this package."""
    if "gsplat" in sys.modules and not force:
        mod = sys.modules["gsplat"]
        if getattr(mod, "__b200__", False):
            return
        raise RuntimeError("a different `gsplat` is already imported; pass force=True to shadow it")
    from . import rendering
    from .cuda import _wrapper
    from . import cuda as _cuda

    pkg = types.ModuleType("gsplat")
    pkg.__b200__ = True
    pkg.__version__ = "1.4.0+b200"
    pkg.__path__ = []  # mark as package
    pkg.rendering = rendering
    pkg.cuda = _cuda
    pkg.rasterization = rendering.rasterization
    pkg.spherical_harmonics = _wrapper.spherical_harmonics
    sys.modules["gsplat"] = pkg
    sys.modules["gsplat.rendering"] = rendering
    sys.modules["gsplat.cuda"] = _cuda
    sys.modules["gsplat.cuda._wrapper"] = _wrapper


def install_as_mtgs_ssim(force: bool = False) -> None:
    """Make ``from mtgs.utils.ssim import MaskedSSIM`` (mtgs/scene_model/mtgs_scene_graph.py:36) resolve to
    ``mtgs_b200.ssim`` without touching the reference tree: the import system consults ``sys.modules`` for the
    fully qualified submodule name before it looks inside the ``mtgs.utils`` package."""
    name = "mtgs.utils.ssim"
    if name in sys.modules and not force and not getattr(sys.modules[name], "__b200__", False):
        raise RuntimeError(f"`{name}` is already imported; pass force=True to shadow it")
    from . import ssim
    ssim.__b200__ = True
    sys.modules[name] = ssim
