"""TEST-ONLY: run a GPU-only script of this repository on the CPU, to catch Python-level mistakes (names, shapes,
argument order, control flow) before the script meets a real GPU:

    python tests/simt/fake_cuda_run.py tools/fit_small.py tiny 3

  * the native kernels run through the SIMT shim (NativeFitLoop / Densifier re-pointed at the emulated library),
  * the msplat operators (gflow_b200.ops) are answered by the CPU oracle,
  * "cuda" devices map to the CPU, torch.cuda.synchronize / streams are no-ops.
Numbers printed by a script run this way mean nothing; only "it ran to the end" does."""
import contextlib
import os
import runpy
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import emu  # noqa: E402
from gflow_b200 import densify, fit, ops  # noqa: E402
from oracle import splat_ref as R  # noqa: E402


def install():
    fit.NativeFitLoop = emu.fit_loop_class()
    densify.Densifier = emu.densifier_class()
    for name in ("project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending", "compute_sh"):
        setattr(ops, name, getattr(R, name))
    ops.rasterization = lambda xyz, scale, rot, op, feat, intr, extr, W, H, bg: R.render_step(xyz, scale, rot, op, feat, intr,
                                                                                             extr, bg, W, H)[0]
    ops.rasterization_unfused = ops.rasterization
    real_device = torch.device

    class FakeDevice:
        def __new__(cls, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                return real_device("cpu")
            return real_device(*a, **k)

    torch.device = FakeDevice
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(real_device("cpu") if isinstance(x, str) and x.startswith("cuda") else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None

    class _Stream:
        def __init__(self, *a, **k):
            pass

        def wait_stream(self, other):
            pass

    torch.cuda.Stream = _Stream
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    # the emulated loop ignores streams
    base = fit.NativeFitLoop

    class NoStream(base):
        def __init__(self, *a, stream=None, **k):
            super().__init__(*a, stream=None, **k)

    fit.NativeFitLoop = NoStream


if __name__ == "__main__":
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
