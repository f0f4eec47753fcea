"""Shared checker of the native fit iteration (csrc/fit.cu) against oracle/fit_ref.py, used with the emulated
kernel library on the CPU (tests/test_simt_fit.py) and with the real library on the GPU
(tests/test_zz_native_fit_gpu.py).

Every iteration is checked at the kernel's own current parameters, so the comparison does not drift:
  * losses of the iteration (mse / SSIM / depth / regularisers) against the oracle's forward,
  * raw-attribute gradients (before masking), dL/d(pose) and the depth_a / depth_b update against autograd,
  * the in-kernel Adam + LinearLR update against torch.optim.Adam fed with the kernel's gradients.
"""
import os

import torch

from conftest import assert_close
from gflow_b200 import fit
from gflow_b200.synthetic import make_scene
from oracle import fit_ref as FR

ATTRS = FR.ATTRS
WIDTH = {"xyz": 3, "scale": 3, "rotate": 4, "opacity": 1, "rgb": 3}


def raw_state(sc, seed):
    g = torch.Generator().manual_seed(seed)
    sign = torch.where(torch.rand(sc.scale.shape, generator=g) < 0.3, -1.0, 1.0)  # abs() must see both signs
    return {"xyz": sc.xyz.clone(), "scale": sc.scale * sign,
            "rotate": sc.rotate * (0.5 + torch.rand(sc.rotate.shape[0], 1, generator=g)),
            "opacity": torch.logit(sc.opacity.clamp(0.02, 0.98)) / 10.0, "rgb": torch.logit(sc.rgb.clamp(0.02, 0.98))}


def make_problem(N=350, W=64, H=48, seed=0):
    sc = make_scene(N, W, H, seed=seed, profile="synthetic")
    raw = raw_state(sc, seed)
    pose = fit.extr_to_pose(sc.extr) * 1.7  # un-normalised quaternion: the normalisation backward matters
    pose[4:] /= 1.7
    # target: the same Gaussians with other colours, seen from a slightly shifted camera
    sc2 = make_scene(N, W, H, seed=seed + 50, profile="synthetic")
    raw2 = dict(raw, rgb=torch.logit(sc2.rgb.clamp(0.02, 0.98)))
    pose2 = pose.clone()
    pose2[4] += 0.03
    with torch.no_grad():
        img, dmap, _, _ = FR.render(raw2, pose2, sc.intr, W, H, 0.0)
    gt_image = img.permute(1, 2, 0).contiguous()
    gt_depth = (dmap.permute(1, 2, 0) * 1.1 + 0.05).contiguous()
    return sc, raw, pose, gt_image, gt_depth


def ref_config(cfg: fit.FitConfig) -> FR.FitRefConfig:
    return FR.FitRefConfig(iterations=cfg.iterations, lr=cfg.lr, lr_camera=cfg.lr_camera, lambda_rgb=cfg.lambda_rgb,
                           use_ssim=cfg.use_ssim, lambda_depth=cfg.lambda_depth, lambda_var=cfg.lambda_var,
                           lambda_scale=cfg.lambda_scale, lambda_still=cfg.lambda_still, lambda_flow=cfg.lambda_flow,
                           camera_only=cfg.camera_only, freeze_rgb=cfg.freeze_rgb, background=cfg.background)


def make_prev(sc, raw, pose, W, H, seed):
    """Previous-frame state for the still / flow terms: the same Gaussians slightly displaced, their projected
    centres, a random still mask and a smooth random flow field."""
    g = torch.Generator().manual_seed(seed + 77)
    N = raw["xyz"].shape[0]
    n = N - 30  # the previous frame had fewer Gaussians (densification appends)
    last_xyz = raw["xyz"][:n] + 0.01 * torch.randn(n, 3, generator=g)
    with torch.no_grad():
        _, _, uv, _ = FR.render(dict(raw, xyz=torch.cat([last_xyz, raw["xyz"][n:]])), pose, sc.intr, W, H, 0.0, want_depth=False)
    gt_flow = 2.0 * torch.randn(H, W, 2, generator=g)
    return dict(last_xyz=last_xyz, last_still_mask=torch.rand(n, generator=g) > 0.4, last_uv=uv[:n].contiguous(),
                gt_flow=gt_flow)


def run_and_check(loop_cls, device, cfg: fit.FitConfig, n_iters, N=350, W=64, H=48, seed=0, pixel_mask=None,
                  still_mask=None, capacity=None, with_prev=False, tentative_still=None):
    """Returns (loop, fitter, raw0, pose0) after `n_iters` checked iterations."""
    sc, raw, pose, gt_image, gt_depth = make_problem(N, W, H, seed)
    prev_ref = prev_dev = None
    if with_prev:
        prev_ref = make_prev(sc, raw, pose, W, H, seed)
        prev_ref["and_mask"] = FR.flow_and_mask(prev_ref["last_uv"], W, H, still_mask, cfg.camera_only)
        assert int(prev_ref["and_mask"].sum()) > 20 and int(prev_ref["last_still_mask"].sum()) > 20
        prev_dev = fit.PrevFrame(**{k: prev_ref[k].to(device) for k in ("last_xyz", "last_still_mask", "last_uv", "gt_flow")})
    use_depth = cfg.lambda_depth > 0
    rcfg = ref_config(cfg)
    dev = torch.device(device)
    fitter = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), pose.to(dev), W, H)
    loop = loop_cls(fitter, gt_image.to(dev), gt_depth.to(dev) if use_depth else None, cfg,
                    pixel_mask=None if pixel_mask is None else pixel_mask.to(dev),
                    still_mask=None if still_mask is None else still_mask.to(dev), capacity=capacity or 40 * N, debug=True,
                    prev=prev_dev, tentative_still=None if tentative_still is None else tentative_still.to(dev))
    dynamic = cfg.camera_only and tentative_still is not None
    keep = None
    if dynamic:
        keep = torch.ones(H, W, dtype=torch.bool) if pixel_mask is None else pixel_mask.clone()
    # shadow optimiser: torch.optim.Adam fed with the KERNEL's gradients
    shadow = {k: raw[k].clone().requires_grad_(True) for k in ATTRS}
    sh_pose = pose.clone().requires_grad_(True)
    sh_ab = torch.tensor([1.0, 0.0], requires_grad=True)
    opt = torch.optim.Adam([{"params": list(shadow.values()), "lr": rcfg.lr}, {"params": [sh_pose], "lr": rcfg.lr_camera},
                            {"params": [sh_ab], "lr": rcfg.lr}])
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.1, total_iters=rcfg.iterations)
    for it in range(n_iters):
        # oracle forward + autograd at the kernel's current parameters
        cur = {k: fitter.attrs[k].data.cpu().clone().requires_grad_(True) for k in ATTRS}
        cur_pose = fitter.pose.data.cpu().clone().requires_grad_(True)
        cur_ab = torch.cat([fitter.depth_a.data.cpu(), fitter.depth_b.data.cpu()]).clone().requires_grad_(True)
        if dynamic:  # the moving subset's footprint under the current pose leaves the losses, cumulatively
            foot = FR.moving_footprint({k: v.detach() for k, v in cur.items()}, cur_pose, sc.intr, W, H, cfg.background,
                                       tentative_still)
            keep = keep & ~foot
            pixel_mask = keep
        loss, parts = FR.iteration_loss(cur, cur_pose, cur_ab, sc.intr, gt_image, gt_depth if use_depth else None,
                                        pixel_mask, W, H, rcfg, prev_ref, still_mask)
        loss.backward()
        assert torch.allclose(loop.camera().cpu()[:12].reshape(3, 4), FR.pose_to_extr(cur_pose.detach()), atol=2e-6)
        loop.run(1)
        h = loop.loss_history()[it].cpu()
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
        if rcfg.lambda_still and with_prev:
            assert abs(float(h[6]) - float(parts["still"])) <= 1e-5 * float(parts["still"]) + 1e-9
        if rcfg.lambda_flow and with_prev:
            assert abs(float(h[7]) - float(parts["flow"])) <= 1e-4 * float(parts["flow"]) + 1e-9
        if dynamic:
            got = loop.pixel_keep_mask()
            got = torch.ones(H, W, dtype=torch.bool) if got is None else got.cpu()  # None: nothing is masked
            assert int((got != keep).sum()) <= 2, "pixel mask carved by the moving subset (grey > 0 is a hard threshold)"
        st = loop.status().cpu()
        assert int(st[0]) == it + 1 and 0 < int(st[1]) <= loop.capacity
        # gradients
        kg = loop.dbg_grads.cpu().clone()
        col = 0
        # with a depth term, frozen colours (frames >= 1 / camera-only) make the blend backward skip the colour
        # channels' own gradient altogether: the reference zeroes it anyway (trainer.py:537-551)
        rgb_skipped = use_depth and (rcfg.freeze_rgb or rcfg.camera_only)
        for k in ATTRS:
            og = cur[k].grad if cur[k].grad is not None else torch.zeros_like(cur[k])
            if k == "rgb" and rgb_skipped:
                assert not kg[:, col:col + WIDTH[k]].any(), "skipped colour gradient must read as zero"
            else:
                assert_close(kg[:, col:col + WIDTH[k]], og.reshape(-1, WIDTH[k]), 1e-3, f"iter {it} grad {k}", outlier_frac=2e-3,
                             outlier_rel=5e-2)
            col += WIDTH[k]
        d_pose = loop.d_pose().cpu().clone()
        assert_close(d_pose, cur_pose.grad, 2e-3, f"iter {it} d_pose")
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
        sh_pose.grad = d_pose
        sh_ab.grad = cur_ab.grad.clone() if use_depth else None
        opt.step()
        sched.step()
        for k in ATTRS:
            assert torch.allclose(fitter.attrs[k].data.cpu(), shadow[k].detach(), rtol=1e-5, atol=2e-6), \
                f"iter {it} Adam update of {k}"
        assert torch.allclose(fitter.pose.data.cpu(), sh_pose.detach(), rtol=1e-5, atol=2e-6), f"iter {it} pose update"
        ab = torch.cat([fitter.depth_a.data.cpu(), fitter.depth_b.data.cpu()])
        assert torch.allclose(ab, sh_ab.detach(), rtol=1e-4, atol=2e-5), f"iter {it} depth_a / depth_b"
    return loop, fitter, raw, pose


def case_list():
    """(name, FitConfig, kwargs) shared by the CPU-emulated and the GPU run."""
    C = fit.FitConfig
    g = torch.Generator().manual_seed(9)
    N, W, H = 350, 64, 48
    pixel_mask = torch.rand(H, W, generator=g) > 0.3
    still = torch.rand(N - 50, generator=g) > 0.5
    return [
        ("mse_depth", C(iterations=10, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True), dict(n_iters=4)),
        ("mse_only_bg", C(iterations=5, lr=1e-2, lambda_depth=0.0, background=0.3, native=True), dict(n_iters=2, seed=2)),
        ("ssim", C(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, native=True),
         dict(n_iters=2, W=70, H=37, seed=3)),
        ("regularisers", C(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, lambda_var=0.7, lambda_scale=0.4,
                           native=True), dict(n_iters=2, seed=4)),
        ("still_flow", C(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, lambda_still=0.5, lambda_flow=0.02,
                         native=True), dict(n_iters=2, seed=7, with_prev=True, still_mask=torch.rand(300, generator=g) > 0.5)),
        ("camera_only", C(iterations=6, lr=4e-3, lr_camera=2e-3, lambda_depth=0.1, camera_only=True, native=True),
         dict(n_iters=3, seed=5)),
        ("camera_only_moving_footprint", C(iterations=6, lr=4e-3, lr_camera=2e-3, lambda_depth=0.1, camera_only=True,
                                           use_ssim=True, native=True),
         dict(n_iters=3, seed=14, tentative_still=torch.rand(330, generator=g) > 0.15, pixel_mask=torch.rand(48, 64, generator=g) > 0.1)),
        ("masks", C(iterations=6, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, freeze_rgb=True, use_ssim=True, lambda_scale=0.3,
                    native=True),
         dict(n_iters=2, N=N, W=W, H=H, seed=6, pixel_mask=pixel_mask, still_mask=still)),
    ]


def post_checks(name, loop, fitter, raw0, pose0, kwargs):
    cur = {k: fitter.attrs[k].data.cpu() for k in ATTRS}
    if name == "mse_depth":
        h = loop.loss_history().cpu()
        assert float(h[-1, 0]) < float(h[0, 0]), "the loss goes down"
    if name == "camera_only_moving_footprint":
        keep = loop.pixel_keep_mask().cpu()
        assert 0 < int(keep.sum()) < int(kwargs["pixel_mask"].sum()), "the moving subset carved pixels out of the static mask"
    if name == "camera_only":
        assert all(torch.equal(cur[k], raw0[k]) for k in ATTRS)
        assert not torch.equal(fitter.pose.data.cpu(), pose0)
    if name == "masks":
        still = kwargs["still_mask"]
        n = still.shape[0]
        assert torch.equal(cur["rgb"], raw0["rgb"])
        assert torch.equal(cur["xyz"][:n][still], raw0["xyz"][:n][still])
        assert not torch.equal(cur["xyz"][:n][~still], raw0["xyz"][:n][~still])


# ----------------------------------------------------------------------------- golden vectors of the reference trainer
GOLDEN_TRAINER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainer_stages.npz")


def load_trainer_golden():
    import numpy as np

    z = np.load(GOLDEN_TRAINER)
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def golden_stage_inputs(G, stage):
    """(FitConfig, start state, pose, targets, masks, prev) of one recorded stage (tests/golden/make_trainer_golden.py)."""
    it, lr, lr_cam, l_rgb, l_depth, l_var, l_scale, l_still, l_flow = (float(v) for v in G[f"{stage}/hyper"])
    later = stage != "first"
    cfg = fit.FitConfig(iterations=int(it), lr=lr, lr_camera=lr_cam, lambda_rgb=l_rgb, lambda_depth=l_depth, lambda_var=l_var,
                        lambda_scale=l_scale, lambda_still=l_still, lambda_flow=l_flow, use_ssim=True, native=True,
                        camera_only=(stage == "camera"), freeze_rgb=later, check_every=int(it))
    raw = {k: G[f"{stage}/state/{k}"][0].clone() for k in ATTRS}
    pose = G[f"{stage}/state/pose"][0].clone()
    img, depth = (G["img1"], G["depth1"]) if later else (G["img0"], G["depth0"])
    kw = dict(pixel_mask=None, still_mask=None, prev=None, tentative_still=None)
    if later:
        kw["still_mask"] = G[f"{stage}/still_mask"].bool()
        kw["prev"] = dict(last_xyz=G[f"{stage}/last_xyz"], last_still_mask=G[f"{stage}/last_still_mask"].bool(),
                          last_uv=G[f"{stage}/last_uv"], gt_flow=G["gt_flow"])
    if stage == "camera":
        kw["pixel_mask"] = ~G["move_mask1"].bool()
        kw["tentative_still"] = G["camera/tentative"].bool()
    return cfg, raw, pose, img, depth, kw


def check_native_stage_against_reference_golden(loop_cls, device, stage, loss_rtol=3e-2, attr_atol=2e-3, attr_frac=0.04,
                                                pose_atol=5e-3):
    """Runs one whole recorded stage with the native loop from the reference's recorded start and compares with what
    the UNMODIFIED reference posted / ended with.  Loss of iteration 0 tight (same state); later iterations and final
    parameters with the slack two differently-rounded Adam trajectories need."""
    G = load_trainer_golden()
    W, H = int(G["W"]), int(G["H"])
    cfg, raw, pose, img, depth, kw = golden_stage_inputs(G, stage)
    dev = torch.device(device)
    f = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, G["intr"].to(dev), pose.to(dev), W, H)
    ab0 = G[f"{stage}/state/ab"][0]
    f.depth_a.data.fill_(float(ab0[0]))
    f.depth_b.data.fill_(float(ab0[1]))
    prev = None if kw["prev"] is None else fit.PrevFrame(**{k: v.to(dev) for k, v in kw["prev"].items()})
    loop = loop_cls(f, img.to(dev), depth.to(dev), cfg, pixel_mask=None if kw["pixel_mask"] is None else kw["pixel_mask"].to(dev),
                    still_mask=None if kw["still_mask"] is None else kw["still_mask"].to(dev), prev=prev,
                    tentative_still=None if kw["tentative_still"] is None else kw["tentative_still"].to(dev), capacity=40 * int(G["N"]))
    loop.run(cfg.iterations)
    ours = loop.loss_history()[:, 0].cpu().double()
    ref = G[f"{stage}/posted/total"].double()
    assert abs(float(ours[0] - ref[0])) <= 3e-4 * abs(float(ref[0])), (stage, ours, ref)
    assert torch.allclose(ours, ref, rtol=loss_rtol), (stage, ours, ref)

    def close(a, b, atol, frac=attr_frac):
        bad = ((a - b).abs() > atol).any(dim=-1)
        return float(bad.float().mean()) <= frac

    for k in ATTRS:
        assert close(f.attrs[k].data.cpu(), G[f"{stage}/final/{k}"], attr_atol), (stage, k)
    assert torch.allclose(f.pose.data.cpu(), G[f"{stage}/final/pose"], atol=pose_atol), stage
    if stage == "camera":
        assert all(torch.equal(f.attrs[k].data.cpu(), raw[k]) for k in ATTRS)
        assert not torch.equal(f.pose.data.cpu(), pose)
    if stage != "first":
        assert torch.equal(f.attrs["rgb"].data.cpu(), raw["rgb"])
    return loop


def check_operator_stage_against_reference_golden(device, stage, loss_rtol=3e-2, attr_atol=2e-3, attr_frac=0.04,
                                                  pose_atol=5e-3):
    """The same recorded stage through the OPERATOR path (FrameFitter.train with native=False: the msplat operators one
    by one + torch autograd + torch.optim.Adam, i.e. the way the unmodified gflow/trainer.py:387-582 drives the drop-in)
    against what the reference posted / ended with."""
    G = load_trainer_golden()
    W, H = int(G["W"]), int(G["H"])
    cfg, raw, pose, img, depth, kw = golden_stage_inputs(G, stage)
    cfg.native = False
    dev = torch.device(device)
    f = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, G["intr"].to(dev), pose.to(dev), W, H)
    ab0 = G[f"{stage}/state/ab"][0]
    f.depth_a.data.fill_(float(ab0[0]))
    f.depth_b.data.fill_(float(ab0[1]))
    prev = None if kw["prev"] is None else fit.PrevFrame(**{k: v.to(dev) for k, v in kw["prev"].items()})
    res = f.train(img.to(dev), depth.to(dev), cfg, pixel_mask=None if kw["pixel_mask"] is None else kw["pixel_mask"].to(dev),
                  still_mask=None if kw["still_mask"] is None else kw["still_mask"].to(dev), prev=prev,
                  tentative_still=None if kw["tentative_still"] is None else kw["tentative_still"].to(dev))
    ours = torch.tensor(res.losses, dtype=torch.float64)
    ref = G[f"{stage}/posted/total"].double()
    assert abs(float(ours[0] - ref[0])) <= 3e-4 * abs(float(ref[0])), (stage, ours, ref)
    assert torch.allclose(ours, ref, rtol=loss_rtol), (stage, ours, ref)
    for k in ATTRS:
        bad = ((f.attrs[k].data.cpu() - G[f"{stage}/final/{k}"]).abs() > attr_atol).any(dim=-1)
        assert float(bad.float().mean()) <= attr_frac, (stage, k, float(bad.float().mean()))
    assert torch.allclose(f.pose.data.cpu(), G[f"{stage}/final/pose"], atol=pose_atol), stage
    if stage == "camera":
        assert all(torch.equal(f.attrs[k].data.cpu(), raw[k]) for k in ATTRS)
    if stage != "first":
        assert torch.equal(f.attrs["rgb"].data.cpu(), raw["rgb"])
    return res

