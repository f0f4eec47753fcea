"""gflow_b200 -- Blackwell-native (sm_100a) differentiable Gaussian-splat rasteriser.

Drop-in for the `msplat` operator surface GFlow calls
(/root/reference/gflow/utils/render.py:21-154).  Importing the package loads the in-tree
CUDA library (building it with nvcc when missing); there is no CPU fallback.
"""
import os as _os
import sys as _sys

from . import capi  # noqa: F401  (fails loudly if libgflow_b200.so cannot be built / loaded)
from .ops import BACKEND  # noqa: F401  "cpp_extension" when the in-tree C++ binding is built, else "ctypes"
from .ops import (  # noqa: F401
    alpha_blending,
    compute_cov3d,
    compute_sh,
    ewa_project,
    project_point,
    rasterization,
    rasterization_unfused,
    sort_gaussian,
)

from .graphs import BatchedRenderStep, GraphedRenderStep  # noqa: F401,E402  (forward + backward as one CUDA graph; a batch of cameras side by side)

__version__ = "1.1"
DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def install_dropin() -> str:
    """Make ``import msplat`` resolve to this implementation (prepends the shim dir to sys.path)."""
    if DROPIN_DIR not in _sys.path:
        _sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR
