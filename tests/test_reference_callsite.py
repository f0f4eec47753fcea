"""The reference's own caller, executed unmodified against our operator signatures (CPU, no GPU).

/root/reference/gflow/utils/render.py is loaded from where it lies (never copied) with `msplat` resolving
to a recording module: every call render.py makes is first bound against the signature of the matching
`gflow_b200.ops` function (positional order, arity, defaults -- the drop-in contract of SURVEY.md 8b),
then answered by the CPU oracle so render.py can run to completion on CPU tensors.  What comes back from
`render_multiple` / `render_traj` is compared with the same chain evaluated directly.

Skipped where /root/reference does not exist (the GPU box); this module is not part of the -m gpu run.
"""
import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

from gflow_b200.synthetic import make_scene
from oracle import splat_ref as R

RENDER_PY = "/root/reference/gflow/utils/render.py"
pytestmark = pytest.mark.skipif(not os.path.exists(RENDER_PY), reason="reference sources not mounted")


def _load_render_with(msplat_module):
    """Import render.py under a throw-away package whose `.color` is a stub (matplotlib is absent)."""
    pkg = types.ModuleType("_ref_utils")
    pkg.__path__ = []
    color = types.ModuleType("_ref_utils.color")

    def apply_float_colormap(image, colormap="turbo", non_zero=False):
        v = image - image[image != 0].min() if non_zero else image - image.min()
        v = (v / (v.max() + 1e-5)).clamp(0, 1)
        return torch.cat([v, 1 - v, 0.5 * v], dim=-1).float()  # (N,3), any smooth map will do

    color.apply_float_colormap = apply_float_colormap
    saved = {k: sys.modules.get(k) for k in ("_ref_utils", "_ref_utils.color", "msplat")}
    sys.modules.update({"_ref_utils": pkg, "_ref_utils.color": color, "msplat": msplat_module})
    try:
        spec = importlib.util.spec_from_file_location("_ref_utils.render", RENDER_PY)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "_ref_utils"
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _recording_msplat():
    """A module named msplat: binds each call against gflow_b200.ops' signature, computes with the oracle."""
    from gflow_b200 import ops

    calls = []
    fake = types.ModuleType("msplat")
    for name in ("project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending"):
        ours = getattr(ops, name + "_py", getattr(ops, name))  # the Python definition carries the signature
        sig = inspect.signature(ours)
        ref_fn = getattr(R, name)

        def make(name=name, sig=sig, ref_fn=ref_fn):
            def call(*args, **kwargs):
                bound = sig.bind(*args, **kwargs)  # TypeError here = our surface does not accept the reference's call
                calls.append((name, len(args), sorted(kwargs)))
                return ref_fn(*bound.args, **bound.kwargs)
            return call

        setattr(fake, name, make())
    return fake, calls


def _input_group(sc):
    return [sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, sc.intr, sc.extr, 0.0, sc.W, sc.H]


def test_render_multiple_runs_unmodified_on_our_signatures():
    fake, calls = _recording_msplat()
    render = _load_render_with(fake)
    sc = make_scene(400, 96, 64, seed=3, profile="gflow")
    out = render.render_multiple(_input_group(sc), ["rgb", "uv", "depth", "depth_map", "depth_map_color", "center"])
    assert [c[0] for c in calls] == ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian"] + ["alpha_blending"] * 4
    assert all(c[2] == [] for c in calls), "render.py calls every operator positionally"
    assert [c[1] for c in calls] == [5, 3, 8, 6, 9, 9, 9, 9]
    assert out["rgb"].shape == (3, sc.H, sc.W) and out["depth_map"].shape == (1, sc.H, sc.W)
    assert out["depth_map_color"].shape == (3, sc.H, sc.W) and out["center"].shape == (3, sc.H, sc.W)
    assert out["uv"].shape == (400, 2) and out["depth"].shape == (400, 1)
    # the same chain evaluated directly
    uv, depth = R.project_point(sc.xyz, sc.intr, sc.extr, sc.W, sc.H)
    vis = depth != 0
    cov = R.compute_cov3d(sc.scale, sc.rotate, vis)
    conic, radius, tiles = R.ewa_project(sc.xyz, cov, sc.intr, sc.extr, uv, sc.W, sc.H, vis)
    ids, rng = R.sort_gaussian(uv, depth, sc.W, sc.H, radius, tiles)
    assert torch.equal(out["uv"], uv) and torch.equal(out["depth"], depth)
    assert torch.allclose(out["rgb"], R.alpha_blending(uv, conic, sc.opacity, sc.rgb, ids, rng, 0.0, sc.W, sc.H))
    assert torch.allclose(out["depth_map"], R.alpha_blending(uv, conic, sc.opacity, depth, ids, rng, 0.0, sc.W, sc.H))
    ident = torch.ones_like(conic) * torch.tensor([1.0, 0.0, 1.0])
    assert torch.allclose(out["center"],
                          R.alpha_blending(uv, ident, torch.ones_like(sc.opacity), sc.rgb, ids, rng, 0.0, sc.W, sc.H))
    # what render.py hands to the sort is what our sort accepts: int32 (N,1) radius / tiles_touched
    assert radius.dtype == torch.int32 and tiles.dtype == torch.int32 and radius.shape == (400, 1)


def test_render_traj_runs_unmodified_on_our_signatures():
    fake, calls = _recording_msplat()
    render = _load_render_with(fake)
    sc = make_scene(300, 80, 48, seed=5, profile="gflow")
    img = render.render_traj(_input_group(sc), point_num=20)
    assert img.shape == (3, sc.H, sc.W) and torch.isfinite(img).all()
    assert [c[0] for c in calls] == ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending"]


def test_both_bindings_expose_the_reference_arity():
    """The C++ binding and the ctypes functions take the same positional parameters the reference uses."""
    from gflow_b200 import ops

    expect = {"project_point": ["xyz", "intr", "extr", "W", "H", "nearest", "extent"],
              "compute_cov3d": ["scale", "rotate", "visible"],
              "ewa_project": ["xyz", "cov3d", "intr", "extr", "uv", "W", "H", "visible"],
              "sort_gaussian": ["uv", "depth", "W", "H", "radius", "tiles_touched"],
              "alpha_blending": ["uv", "conic", "opacity", "feature", "gaussian_ids_sorted", "tile_range", "bg", "W", "H",
                                 "ndc"],
              "compute_sh": ["shs", "dirs", "visible"]}
    for name, params in expect.items():
        for fn in {getattr(ops, name), getattr(ops, name + "_py")}:
            assert list(inspect.signature(fn).parameters) == params, name
