"""Shared harness that imports the UNMODIFIED /root/reference/gflow/trainer.py on the CPU (used by
tests/test_reference_trainer.py and tests/golden/make_trainer_golden.py):
  * `msplat` -> a module that binds every call against the signature of the matching gflow_b200.ops function (the
    drop-in contract) and computes with the CPU oracle (oracle/splat_ref.py),
  * absent third-party packages -> tests/shims,  * Tensor.cuda() -> identity,  * tqdm -> a recorder of the posted losses,
  * the post-stage concave hull (shapely; outside the path) -> an empty mask.
Nothing here is imported by gflow_b200/."""
import contextlib
import inspect
import os
import sys
import types

import numpy as np
import torch

from oracle import splat_ref as R

REF = "/root/reference/gflow"
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.exists(os.path.join(REF, "trainer.py"))


def oracle_msplat(calls):
    from gflow_b200 import ops

    fake = types.ModuleType("msplat")

    def wrap(name):
        sig = inspect.signature(getattr(ops, name))
        impl = getattr(R, name)

        def f(*a, **k):
            sig.bind(*a, **k)  # raises TypeError if the reference's call does not fit our signature
            calls.append(name)
            return impl(*a, **k)

        return f

    for n in ("project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending", "compute_sh"):
        setattr(fake, n, wrap(n))
    return fake


class Bar:
    """tqdm stand-in that records the loss dictionaries the reference posts every iteration (trainer.py:556-557)."""
    posted = []

    def __init__(self, *a, **k):
        pass

    def set_postfix(self, d):
        Bar.posted.append(dict(d))

    def update(self, n=1):
        pass

    def close(self):
        pass


class _Hull:
    def __init__(self, pts, *a, **k):
        pass

    def mask(self, w, h):
        return np.zeros((h, w), dtype=np.float32)


@contextlib.contextmanager
def reference_trainer(workdir):
    """Yields (trainer module, list of msplat calls made)."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import shims

    calls = []
    added = shims.install()
    saved = {k: sys.modules.get(k) for k in ("msplat", "utils", "trainer")}
    sys.modules["msplat"] = oracle_msplat(calls)
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(k)
    sys.path.insert(0, REF)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    cwd = os.getcwd()
    os.makedirs(str(workdir), exist_ok=True)
    os.chdir(str(workdir))
    try:
        import trainer as ref_trainer  # noqa: the reference module, unmodified
        import utils as ref_utils

        ref_trainer.tqdm = Bar
        ref_utils.FastConcaveHull2D = _Hull
        Bar.posted = []
        yield ref_trainer, calls
    finally:
        os.chdir(cwd)
        torch.Tensor.cuda = orig_cuda
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k in ("trainer", "utils") or k.startswith("utils.")] + added:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)


def scene(W=48, H=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    img = torch.stack([0.5 + 0.5 * torch.sin(9 * xx + 3 * yy), 0.5 + 0.5 * torch.cos(7 * yy), xx * yy], dim=-1)
    img = (img + 0.05 * torch.rand(H, W, 3, generator=g)).clamp(0.02, 0.98).float()
    depth = (1.5 + xx + 0.5 * torch.sin(5 * yy)).unsqueeze(-1).float()
    return img, depth


def new_trainer(ref_trainer, workdir, W=48, H=32, N=300):
    img, depth = scene(W, H)
    np.random.seed(0)
    torch.manual_seed(0)
    t = ref_trainer.SimpleGaussian(gt_image=img, gt_depth=depth, num_points=N, sequence_path=os.path.join(str(workdir), "seq"))
    t.load_camera(focal=0.6 * W, pp=[W / 2.0, H / 2.0], show=False)
    t.init_gaussians_from_image(gt_image=img, gt_depth=depth, num_points=N)
    return t, img, depth
