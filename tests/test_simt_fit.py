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


def test_densify_between_iterations_recreates_the_optimiser_like_the_reference():
    """trainer.py:566-571,941-951: new Gaussians are appended, Adam restarts over the attributes only with the
    initial lr; pose and depth_a / depth_b stop moving.  Checked against torch.optim.Adam fed with the
    kernel's gradients, as in fit_check."""
    from oracle import fit_ref as FR

    N, W, H = 350, 64, 48
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N, W, H, seed=12)
    cfg = fit.FitConfig(iterations=8, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True, check_every=1, num_points=N)
    f = fit.FrameFitter(raw, sc.intr, pose, W, H)
    loop = emu.fit_loop_class()(f, gt_image, gt_depth, cfg, capacity=40 * N, debug=True)
    loop.run(2)
    pose_before, ab_before = f.pose.data.clone(), loop.depth_ab.clone()
    hist_before = loop.loss_history().clone()
    old = {k: f.attrs[k].data.clone() for k in fit.ATTRS}
    uv_before, depth_before = loop.last_uv(), loop.last_depth()
    added = loop.densify(error_threshold=1e-5, percent=0.5, seed=3)
    assert added > 10 and loop.N == N + added
    # a stage may END with a densification (SimpleGaussian.train's defaults: iterations = densify_interval = 500): the
    # caller then builds still masks / last_uv from last_uv() / last_depth(), which must not be fresh-workspace garbage
    uv_after, depth_after = loop.last_uv(), loop.last_depth()
    assert uv_after.shape == (N + added, 2) and torch.equal(uv_after[:N], uv_before) and torch.equal(depth_after[:N], depth_before)
    assert bool(torch.isfinite(uv_after).all()) and bool((depth_after[N:] > 0).all())
    for k in fit.ATTRS:
        assert f.attrs[k].shape[0] == N + added and torch.equal(f.attrs[k].data[:N], old[k])
    assert torch.equal(loop.loss_history(), hist_before) and int(loop.status()[0]) == 2
    # the appended Gaussians project onto the pixels they were drawn from, at the prior's depth
    with torch.no_grad():
        _, _, uv, depth = FR.render({k: f.attrs[k].data for k in fit.ATTRS}, f.pose.data, sc.intr, W, H, 0.0, want_depth=False)
    new_uv = uv[N:]
    seen = depth[N:, 0] != 0
    assert bool(seen.float().mean() > 0.9)
    # geometry.py:115 back-projects with fx on BOTH axes (GFlow's cameras have fx = fy; this one has 32 / 24), so
    # u lands on the drawn pixel column and v on cy + (row - cy) fy / fx
    pix = loop._last_densify_pixels.long()[seen]
    col, row = (pix % W).float(), (pix // W).float()
    fx, fy, cx, cy = (float(v) for v in sc.intr)
    assert torch.allclose(new_uv[seen, 0], col, atol=2e-3)
    assert torch.allclose(new_uv[seen, 1], cy + (row - cy) * fy / fx, atol=2e-3)
    assert torch.allclose(uv_after[N:][seen], new_uv[seen], atol=2e-3), "the carried-over uv of the new rows = their projection"
    assert torch.allclose(depth[N:][seen, 0], gt_depth.reshape(-1)[pix], rtol=1e-4)
    # next iterations: fresh Adam (t restarts at 1), constant lr, camera frozen
    shadow = {k: f.attrs[k].data.clone().requires_grad_(True) for k in fit.ATTRS}
    opt = torch.optim.Adam(list(shadow.values()), lr=cfg.lr)
    for it in range(2):
        loop.run(1)
        kg, col = loop.dbg_grads.clone(), 0
        for k in fit.ATTRS:
            w = shadow[k].shape[1]
            shadow[k].grad = kg[:, col:col + w].clone()
            col += w
        opt.step()
        for k in fit.ATTRS:
            assert torch.allclose(f.attrs[k].data, shadow[k].detach(), rtol=1e-5, atol=2e-6), (it, k)
        assert torch.equal(f.pose.data, pose_before) and torch.equal(loop.depth_ab, ab_before)
    assert loop.done == 4 and int(loop.status()[0]) == 4


def test_train_native_with_densification_schedule(monkeypatch):
    monkeypatch.setattr(fit, "NativeFitLoop", emu.fit_loop_class())
    monkeypatch.setattr(fit.FrameFitter, "render", lambda self, bg=0.0, want_depth=True, with_depth=False: (None, None, None))
    N, W, H = 300, 64, 48
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N, W, H, seed=13)
    f = fit.FrameFitter(raw, sc.intr, pose, W, H)
    cfg = fit.FitConfig(iterations=7, lr=4e-3, lambda_depth=0.1, native=True, densify_interval=2, densify_times=2,
                        densify_err_thre=1e-5, densify_err_percent=0.3)
    res = f.train(gt_image, gt_depth, cfg)
    assert len(res.losses) == 7 and all(v == v for v in res.losses)
    n_final = f.attrs["xyz"].shape[0]
    assert n_final > N and all(f.attrs[k].shape[0] == n_final for k in fit.ATTRS)


def test_occlusion_densification_after_iteration_zero(monkeypatch):
    """trainer.py:562-564: on a later frame, Gaussians are drawn uniformly inside the occlusion mask right after
    iteration 0 and the optimiser is re-created."""
    monkeypatch.setattr(fit, "NativeFitLoop", emu.fit_loop_class())
    monkeypatch.setattr(fit.FrameFitter, "render", lambda self, bg=0.0, want_depth=True, with_depth=False: (None, None, None))
    N, W, H = 300, 64, 48
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N, W, H, seed=15)
    prev = fit.PrevFrame(**fit_check.make_prev(sc, raw, pose, W, H, 15))
    occ = torch.zeros(H, W, 1)
    occ[10:30, 20:50] = 1.0
    f = fit.FrameFitter(raw, sc.intr, pose, W, H)
    cfg = fit.FitConfig(iterations=4, lr=4e-3, lambda_depth=0.1, lambda_flow=0.01, native=True, densify_occ_percent=0.5)
    res = f.train(gt_image, gt_depth, cfg, prev=prev, occlusion_mask=occ, still_mask=torch.rand(270, generator=torch.Generator().manual_seed(0)) > 0.5)
    added = f.attrs["xyz"].shape[0] - N
    assert added == int(N * (600 / (W * H)) * 0.5) and len(res.losses) == 4
    # the new Gaussians sit inside the mask: project them with the oracle
    from oracle import fit_ref as FR

    with torch.no_grad():
        _, _, uv, depth = FR.render({k: f.attrs[k].data for k in fit.ATTRS}, f.pose.data, sc.intr, W, H, 0.0, want_depth=False)
    u = uv[N:, 0]
    assert bool(((u > 19.4) & (u < 49.6)).all())


def test_frames_side_by_side_equal_frames_one_after_the_other():
    """fit_frames_concurrently: round-robin enqueue / check over several loops (streams are a no-op in the emulation);
    each frame's result is bit-identical to running it alone, including a roll-back in one of them."""
    Loop = emu.fit_loop_class()
    cfg = fit.FitConfig(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True, check_every=2)
    probs = [fit_check.make_problem(N=200 + 31 * i, W=64, H=48, seed=30 + i) for i in range(3)]

    def fitters():
        return [fit.FrameFitter(raw, sc.intr, pose, 64, 48) for sc, raw, pose, _, _ in probs]

    alone = fitters()
    for f, (_, _, _, gi, gd) in zip(alone, probs):
        Loop(f, gi, gd, cfg, capacity=40 * 300).run(5)
    together = fitters()

    class SmallFirst(Loop):  # the first workspace is too small: that loop rolls back while the others go on
        def __init__(self, fitter, gi, gd, cfg, stream=None):
            super().__init__(fitter, gi, gd, cfg, capacity=150 if fitter is together[1] else 40 * 300, stream=stream)

    loops = fit.fit_frames_concurrently(together, [(p[3], p[4]) for p in probs], cfg, streams=[None, None, None], loop_cls=SmallFirst)
    assert loops[1].capacity > 150 and all(lp.done == 5 for lp in loops)
    for a, b in zip(alone, together):
        for k in fit.ATTRS:
            assert torch.equal(a.attrs[k].data, b.attrs[k].data)
        assert torch.equal(a.pose.data, b.pose.data) and torch.equal(a.depth_a.data, b.depth_a.data)


def test_empty_moving_subset_with_a_bright_background_masks_everything():
    """Reference quirk kept: with no tentatively-moving Gaussian the subset render is the bare background, grey > 0 for
    any background > 0, so every pixel leaves the losses (trainer.py:446-451)."""
    cfg = fit.FitConfig(iterations=3, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, camera_only=True, background=0.4, native=True)
    fit_check.run_and_check(emu.fit_loop_class(), "cpu", cfg, n_iters=2, N=60, W=33, H=31, seed=3,
                            tentative_still=torch.ones(40, dtype=torch.bool))
