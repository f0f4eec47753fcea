"""Drop-in replacement for the `msplat` package, backed by gflow_b200 (hand-written sm_100a CUDA).

Put ``<repo>/gflow_b200/dropin`` on PYTHONPATH (or call ``gflow_b200.install_dropin()``) and the
unmodified GFlow sources (``import msplat`` at /root/reference/gflow/trainer.py:7 and
/root/reference/gflow/utils/render.py:2) resolve to this module.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from gflow_b200.ops import (  # noqa: E402,F401
    alpha_blending,
    compute_cov3d,
    compute_sh,
    ewa_project,
    project_point,
    rasterization,
    sort_gaussian,
)

__all__ = ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending", "compute_sh",
           "rasterization"]
__version__ = "gflow_b200-1.0"

# opt-in: run the optimisation stages of GFlow's trainer in the native loop (gflow_b200/accelerate.py)
if _os.environ.get("GFLOW_B200_NATIVE_TRAIN") == "1":
    from gflow_b200 import accelerate as _accelerate

    _accelerate.install_import_hook("trainer")

    # `trainer` is normally half-way through its own import when it imports this module (trainer.py:7), so the hook
    # cannot patch it yet; besides the import watcher, the operators look once more the first time one is called.
    def _patch_on_first_call(fn):
        def first(*a, **k):
            _accelerate.try_patch_loaded("trainer")
            for name, raw in _RAW.items():  # from now on the plain operators again
                globals()[name] = raw
            return fn(*a, **k)

        first.__name__, first.__doc__ = fn.__name__, fn.__doc__
        return first

    _RAW = {n: globals()[n] for n in __all__}
    for _n, _f in _RAW.items():
        globals()[_n] = _patch_on_first_call(_f)
