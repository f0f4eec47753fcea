"""The native per-frame optimisation iteration (gflow_b200/csrc/fit.cu) executed on the CPU through the SIMT
shim (tests/simt/) against the PyTorch-autograd restatement of GFlow's loop (oracle/fit_ref.py).

Every iteration is checked at the kernel's own current parameters, so the comparison does not drift:
  * losses of the iteration (mse / SSIM / depth / regularisers) against the oracle's forward,
  * raw-attribute gradients (before masking), dL/d(pose) and the depth_a / depth_b update against autograd,
  * the in-kernel Adam + LinearLR update against torch.optim.Adam fed with the kernel's gradients.
"""
import os
import sys

import pytest
import torch

from conftest import assert_close
from gflow_b200.synthetic import make_scene
from oracle import fit_ref as FR

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import emu  # noqa: E402

ATTRS = FR.ATTRS
WIDTH = {"xyz": 3, "scale": 3, "rotate": 4, "opacity": 1, "rgb": 3}


def _raw_state(sc, seed):
    g = torch.Generator().manual_seed(seed)
    sign = torch.where(torch.rand(sc.scale.shape, generator=g) < 0.3, -1.0, 1.0)  # abs() must see both signs
    return {"xyz": sc.xyz.clone(), "scale": sc.scale * sign, "rotate": sc.rotate * (0.5 + torch.rand(sc.rotate.shape[0], 1, generator=g)),
            "opacity": torch.logit(sc.opacity.clamp(0.02, 0.98)) / 10.0, "rgb": torch.logit(sc.rgb.clamp(0.02, 0.98))}


def _extr_to_pose(extr):
    from gflow_b200.fit import extr_to_pose

    return extr_to_pose(extr)


def _problem(N=350, W=64, H=48, seed=0):
    sc = make_scene(N, W, H, seed=seed, profile="synthetic")
    raw = _raw_state(sc, seed)
    pose = _extr_to_pose(sc.extr) * 1.7  # un-normalised quaternion: the normalisation backward matters
    pose[4:] /= 1.7
    # target: the same scene seen from a slightly different camera with different colours
    sc2 = make_scene(N, W, H, seed=seed + 50, profile="synthetic")
    raw2 = dict(raw, rgb=torch.logit(sc2.rgb.clamp(0.02, 0.98)))
    pose2 = pose.clone()
    pose2[4] += 0.03
    with torch.no_grad():
        img, dmap, _, _ = FR.render(raw2, pose2, sc.intr, W, H, 0.0)
    gt_image = img.permute(1, 2, 0).contiguous()
    gt_depth = (dmap.permute(1, 2, 0) * 1.1 + 0.05).contiguous()
    return sc, raw, pose, gt_image, gt_depth


def _cat_grads(g):
    return torch.cat([g[k].reshape(-1, WIDTH[k]) for k in ATTRS], dim=1)


def _run_and_check(cfg, n_iters, N=350, W=64, H=48, seed=0, pixel_mask=None, still_mask=None, capacity=None):
    sc, raw, pose, gt_image, gt_depth = _problem(N, W, H, seed)
    use_depth = cfg.get("lambda_depth", 0.0) > 0
    rcfg = FR.FitRefConfig(iterations=cfg["iterations"], lr=cfg["lr"], lr_camera=cfg.get("lr_camera", 0.0),
                           lambda_rgb=cfg.get("lambda_rgb", 1.0), use_ssim=cfg.get("use_ssim", False),
                           lambda_depth=cfg.get("lambda_depth", 0.0), lambda_var=cfg.get("lambda_var", 0.0),
                           lambda_scale=cfg.get("lambda_scale", 0.0), camera_only=cfg.get("camera_only", False),
                           freeze_rgb=cfg.get("freeze_rgb", False), background=cfg.get("background", 0.0))
    nf = emu.NativeFit(raw, pose, sc.intr, gt_image, gt_depth if use_depth else None, W, H, cfg,
                       capacity=capacity or 40 * N, pixel_mask=pixel_mask, still_mask=still_mask)
    # shadow optimiser: torch.optim.Adam fed with the KERNEL's gradients
    shadow = {k: raw[k].clone().requires_grad_(True) for k in ATTRS}
    sh_pose = pose.clone().requires_grad_(True)
    sh_ab = torch.tensor([1.0, 0.0], requires_grad=True)
    opt = torch.optim.Adam([{"params": list(shadow.values()), "lr": rcfg.lr}, {"params": [sh_pose], "lr": rcfg.lr_camera},
                            {"params": [sh_ab], "lr": rcfg.lr}])
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.1, total_iters=rcfg.iterations)
    for it in range(n_iters):
        # oracle forward + autograd at the kernel's current parameters
        cur = {k: nf.raw[k].clone().requires_grad_(True) for k in ATTRS}
        cur_pose = nf.pose.clone().requires_grad_(True)
        cur_ab = nf.depth_ab.clone().requires_grad_(True)
        loss, parts = FR.iteration_loss(cur, cur_pose, cur_ab, sc.intr, gt_image, gt_depth if use_depth else None,
                                        pixel_mask, W, H, rcfg)
        loss.backward()
        assert torch.allclose(nf.cam()[:12].reshape(3, 4), FR.pose_to_extr(nf.pose), atol=2e-6), "camera of the iteration"
        nf.iterate(1)
        h = nf.loss_hist()[it]
        assert abs(float(h[0]) - float(parts["total"])) <= 2e-4 * max(1.0, abs(float(parts["total"]))), (it, h, parts)
        assert abs(float(h[1]) - float(parts["mse"])) <= 2e-4 * float(parts["mse"]) + 1e-7
        if rcfg.use_ssim:
            assert abs(float(h[2]) - float(parts["ssim"])) <= 2e-4
        if use_depth:
            assert abs(float(h[3]) - float(parts["depth"])) <= 5e-4 * float(parts["depth"]) + 1e-7
        if rcfg.lambda_var:
            assert abs(float(h[4]) - float(parts["var"])) <= 1e-5 * float(parts["var"]) + 1e-9
        if rcfg.lambda_scale:
            assert abs(float(h[5]) - float(parts["scale"])) <= 1e-5 * float(parts["scale"]) + 1e-9
        assert int(nf.status[0]) == it + 1 and 0 < int(nf.status[1]) <= nf.cap
        # gradients
        kg = nf.dbg_grads.clone()
        og = _cat_grads({k: cur[k].grad if cur[k].grad is not None else torch.zeros_like(cur[k]) for k in ATTRS})
        col = 0
        for k in ATTRS:
            assert_close(kg[:, col:col + WIDTH[k]], og[:, col:col + WIDTH[k]], 1e-3, f"iter {it} grad {k}", outlier_frac=2e-3,
                         outlier_rel=5e-2)
            col += WIDTH[k]
        assert_close(nf.d_pose.clone(), cur_pose.grad, 2e-3, f"iter {it} d_pose")
        # the kernel's update == torch Adam fed with the kernel's gradients and the reference's masks
        col = 0
        for k in ATTRS:
            g = kg[:, col:col + WIDTH[k]].reshape(shadow[k].shape).clone()
            col += WIDTH[k]
            if rcfg.camera_only or (k == "rgb" and rcfg.freeze_rgb):
                g.zero_()
            if k == "xyz" and still_mask is not None:
                g[: still_mask.shape[0]][still_mask] = 0.0
            shadow[k].grad = g
        sh_pose.grad = nf.d_pose.clone()
        sh_ab.grad = cur_ab.grad.clone() if use_depth else None
        opt.step()
        sched.step()
        for k in ATTRS:
            assert torch.allclose(nf.raw[k], shadow[k].detach(), rtol=1e-5, atol=2e-6), f"iter {it} Adam update of {k}"
        assert torch.allclose(nf.pose, sh_pose.detach(), rtol=1e-5, atol=2e-6), f"iter {it} Adam update of the pose"
        assert torch.allclose(nf.depth_ab, sh_ab.detach(), rtol=1e-4, atol=2e-5), f"iter {it} depth_a / depth_b"
    return nf


def test_mse_and_depth_loop():
    nf = _run_and_check(dict(iterations=10, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1), n_iters=4)
    h = nf.loss_hist()
    assert float(h[3, 0]) < float(h[0, 0]), "the loss goes down"


def test_mse_only_three_channel_blend():
    _run_and_check(dict(iterations=5, lr=1e-2, lr_camera=0.0, lambda_depth=0.0, background=0.3), n_iters=2, seed=2)


def test_ssim_term():
    _run_and_check(dict(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True), n_iters=2, W=70, H=37, seed=3)


def test_regularisers():
    _run_and_check(dict(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, lambda_var=0.7, lambda_scale=0.4), n_iters=2,
                   seed=4)


def test_camera_only_freezes_attributes():
    nf = _run_and_check(dict(iterations=6, lr=4e-3, lr_camera=2e-3, lambda_depth=0.1, camera_only=True), n_iters=3, seed=5)
    sc, raw, pose, _, _ = _problem(seed=5)
    assert all(torch.equal(nf.raw[k], raw[k]) for k in ATTRS)
    assert not torch.equal(nf.pose, pose)


def test_masks():
    N, W, H = 350, 64, 48
    g = torch.Generator().manual_seed(9)
    pixel_mask = torch.rand(H, W, generator=g) > 0.3
    still = torch.rand(N - 50, generator=g) > 0.5
    nf = _run_and_check(dict(iterations=6, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, freeze_rgb=True, use_ssim=True), n_iters=2,
                        N=N, W=W, H=H, seed=6, pixel_mask=pixel_mask, still_mask=still)
    _, raw, _, _, _ = _problem(N, W, H, seed=6)
    assert torch.equal(nf.raw["rgb"], raw["rgb"])
    assert torch.equal(nf.raw["xyz"][: N - 50][still], raw["xyz"][: N - 50][still])
    assert not torch.equal(nf.raw["xyz"][: N - 50][~still], raw["xyz"][: N - 50][~still])


def test_capacity_overflow_is_visible_in_status():
    cfg = dict(iterations=3, lr=4e-3, lambda_depth=0.1)
    sc, raw, pose, gt_image, gt_depth = _problem()
    nf = emu.NativeFit(raw, pose, sc.intr, gt_image, gt_depth, 64, 48, cfg, capacity=200)
    nf.iterate(1)
    assert int(nf.status[2]) > 200  # max K seen > capacity: the caller has to redo the chunk with a larger workspace


def test_bad_arguments():
    from gflow_b200.capi import FitLayout

    L = emu.load()
    lay = FitLayout()
    import ctypes

    assert L.gfb_fit_get_layout(0, 64, 48, 100, 10, ctypes.addressof(lay)) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 0, ctypes.addressof(lay)) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 10, None) == -1
    assert L.gfb_fit_get_layout(10, 64, 48, 100, 10, ctypes.addressof(lay)) == 0 and lay.total > 0
    assert L.gfb_fit_iterate(None, None, 100, 10, 0, 1, None) == -1
