"""The native per-frame optimisation iteration (gflow_b200/csrc/fit.cu) executed on the CPU through the SIMT
shim (tests/simt/), driven by the product's own host class (gflow_b200.fit.NativeFitLoop re-pointed at the
emulated library), against the PyTorch-autograd restatement of GFlow's loop (oracle/fit_ref.py).
See tests/fit_check.py for what is compared."""
import ctypes
import os
import sys

import pytest
import torch

import fit_check
from gflow_b200 import fit

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import emu  # noqa: E402

CASES = fit_check.case_list()


@pytest.mark.parametrize("name,cfg,kwargs", CASES, ids=[c[0] for c in CASES])
def test_native_iteration_matches_oracle(name, cfg, kwargs):
    loop, fitter, raw0, pose0 = fit_check.run_and_check(emu.fit_loop_class(), "cpu", cfg, **kwargs)
    fit_check.post_checks(name, loop, fitter, raw0, pose0, kwargs)


def test_overflowing_chunk_is_rolled_back_and_redone():
    """A workspace that is too small truncates tiles; the host loop must notice (max K in the status block),
    restore the parameters / Adam state of the chunk start, grow the workspace and redo the chunk."""
    cfg = fit.FitConfig(iterations=6, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True, check_every=2)
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem()
    Loop = emu.fit_loop_class()

    def run(capacity):
        f = fit.FrameFitter(raw, sc.intr, pose, 64, 48)
        loop = Loop(f, gt_image, gt_depth, cfg, capacity=capacity)
        loop.run(6)
        return f, loop

    f_small, loop_small = run(200)
    f_big, loop_big = run(40 * 350)
    assert loop_small.capacity > 200 and int(loop_small.status()[2]) <= loop_small.capacity
    assert loop_small.done == 6 and int(loop_small.status()[0]) == 6
    for k in fit.ATTRS:
        assert torch.equal(f_small.attrs[k].data, f_big.attrs[k].data), k
    assert torch.equal(f_small.pose.data, f_big.pose.data)
    assert torch.equal(loop_small.loss_history(), loop_big.loss_history())
    assert torch.equal(f_small.depth_a.data, f_big.depth_a.data)


def test_train_native_entry_point_runs_the_loop(monkeypatch):
    """FrameFitter.train(cfg.native=True) drives NativeFitLoop; on CPU tensors the product class refuses."""
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem()
    f = fit.FrameFitter(raw, sc.intr, pose, 64, 48)
    cfg = fit.FitConfig(iterations=3, native=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.train(gt_image, gt_depth, cfg)


def test_bad_arguments():
    from gflow_b200.capi import FitLayout

    L = emu.load()
    lay = FitLayout()
    assert L.gfb_fit_get_layout(0, 64, 48, 100, 10, ctypes.addressof(lay)) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 0, ctypes.addressof(lay)) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 10, None) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 10, ctypes.addressof(lay)) == 0 and lay.total > 0
    assert L.gfb_fit_iterate(None, None, 100, 10, 0, 1, None) == -1


def test_train_native_with_every_term_through_the_emulated_loop(monkeypatch):
    """FrameFitter.train(native=True) end to end on the host side (what tools/fit_small.py does on the GPU):
    every loss term on, masks, previous-frame state, chunked run with a workspace that has to grow."""
    monkeypatch.setattr(fit, "NativeFitLoop", emu.fit_loop_class())
    monkeypatch.setattr(fit.FrameFitter, "render", lambda self, bg=0.0, want_depth=True, with_depth=False: (None, None, None))
    N, W, H = 351, 64, 48  # odd N
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N, W, H, seed=8)
    g = torch.Generator().manual_seed(1)
    prev_ref = fit_check.make_prev(sc, raw, pose, W, H, 8)
    prev = fit.PrevFrame(**prev_ref)
    f = fit.FrameFitter(raw, sc.intr, pose, W, H)
    cfg = fit.FitConfig(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, lambda_var=0.1,
                        lambda_scale=0.1, lambda_still=0.1, lambda_flow=0.01, native=True, check_every=2)
    res = f.train(gt_image, gt_depth, cfg, pixel_mask=torch.rand(H, W, generator=g) > 0.1,
                  still_mask=torch.rand(N - 30, generator=g) > 0.5, prev=prev)
    assert len(res.losses) == 5 and all(v == v and v > 0 for v in res.losses)
    assert res.losses[-1] < res.losses[0]
    assert not torch.equal(f.depth_a.data, torch.ones(1)), "depth_a is copied back from the device-side pair"
