"""The reference's own training loop, executed UNMODIFIED on the CPU, against this repository's restatements.

/root/reference/gflow/trainer.py is imported from where it lies (never copied) with
  * `msplat` resolving to a module that binds every call against the signature of the matching
    gflow_b200.ops function (the drop-in contract) and computes with the CPU oracle (oracle/splat_ref.py),
  * the absent third-party packages replaced by tests/shims (roma, imageio, matplotlib, shapely, concave_hull),
  * Tensor.cuda() made the identity (the reference hard-codes .cuda(); there is no GPU here).
SimpleGaussian.train then runs a few real iterations.  Checked:
  * the run completes through the msplat surface (project_point, compute_cov3d, ewa_project, sort_gaussian,
    alpha_blending x4 per iteration, autograd through all of them),
  * the losses the reference reports for its first iteration equal oracle/fit_ref.iteration_loss at the same
    parameters -- this pins fit_ref (the oracle of the native loop, csrc/fit.cu) to the reference's own code,
  * the parameters after k iterations equal oracle/fit_ref.fit_loop from the same start (Adam, LinearLR).
Skipped where /root/reference is not mounted (the GPU box).
"""
import inspect
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import fit_ref as FR
from oracle import splat_ref as R

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/gflow/trainer.py"), reason="reference sources not mounted")


import ref_harness
from ref_harness import Bar as _Bar
from ref_harness import scene as _scene


@pytest.fixture()
def reference_trainer(tmp_path):
    with ref_harness.reference_trainer(tmp_path) as (mod, calls):
        yield mod, calls


def test_first_frame_stage_runs_unmodified_and_matches_the_fit_oracle(reference_trainer, tmp_path):
    ref_trainer, calls = reference_trainer
    W, H, N, iters = 48, 32, 300, 3
    img, depth = _scene(W, H)
    np.random.seed(0)
    torch.manual_seed(0)
    t = ref_trainer.SimpleGaussian(gt_image=img, gt_depth=depth, num_points=N, sequence_path=str(tmp_path / "seq"))
    t.load_camera(focal=0.6 * W, pp=[W / 2.0, H / 2.0], show=False)
    t.init_gaussians_from_image(gt_image=img, gt_depth=depth, num_points=N)
    start = {k: v.detach().clone() for k, v in t._attributes.items()}
    pose0 = t.pose.detach().clone()
    move_mask = torch.zeros(H, W, dtype=torch.bool)
    move_mask[10:20, 5:25] = True
    lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.2, lambda_scale=0.05)
    t.train(iterations=iters, lr=4e-3, lr_camera=1e-3, move_mask=move_mask, densify_interval=500, densify_times=0, **lam)
    # --- the run went through the whole operator surface
    per_iter = ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian"] + ["alpha_blending"] * 4
    assert calls[: len(per_iter) * iters] == per_iter * iters
    assert len(_Bar.posted) == iters
    # --- first-iteration losses == oracle/fit_ref at the same parameters (depth_den_min=0: the reference does not clamp)
    cfg = FR.FitRefConfig(iterations=iters, lr=4e-3, lr_camera=1e-3, use_ssim=True, depth_den_min=0.0, **lam)
    loss, parts = FR.iteration_loss(start, pose0, torch.tensor([1.0, 0.0]), t.intr, img, depth, None, W, H, cfg)
    p0 = _Bar.posted[0]
    assert abs(float(p0["total"]) - float(parts["total"])) <= 1e-5 * abs(float(parts["total"]))
    assert abs(float(p0["rgb"]) - float(parts["mse"] + 1 - parts["ssim"])) <= 2e-6
    assert abs(float(p0["depth"]) - float(parts["depth"])) <= 2e-6
    assert abs(float(p0["var"]) - float(parts["var"])) <= 2e-6
    assert abs(float(p0["scale"]) - float(parts["scale"])) <= 2e-6
    # --- k iterations of the reference == k iterations of fit_ref.fit_loop (Adam + LinearLR, pose, depth_a/b)
    raw, pose, ab, hist = FR.fit_loop(start, pose0, t.intr, img, depth, W, H, cfg)
    for i in range(iters):
        assert abs(float(_Bar.posted[i]["total"]) - float(hist[i]["total"])) <= 1e-4 * abs(float(hist[i]["total"])), i
    for k in FR.ATTRS:
        assert torch.allclose(t._attributes[k].detach(), raw[k], rtol=1e-4, atol=1e-5), k
    assert torch.allclose(t.pose.detach(), pose, rtol=1e-4, atol=1e-6)
    assert torch.allclose(torch.cat([t.depth_a.detach(), t.depth_b.detach()]), ab, rtol=1e-4, atol=1e-6)
    assert float(_Bar.posted[-1]["total"]) < float(_Bar.posted[0]["total"])


def _first_frame(ref_trainer, tmp_path, W=48, H=32, N=300):
    img, depth = _scene(W, H)
    np.random.seed(0)
    torch.manual_seed(0)
    t = ref_trainer.SimpleGaussian(gt_image=img, gt_depth=depth, num_points=N, sequence_path=str(tmp_path / "seq"))
    t.load_camera(focal=0.6 * W, pp=[W / 2.0, H / 2.0], show=False)
    t.init_gaussians_from_image(gt_image=img, gt_depth=depth, num_points=N)
    move_mask = torch.zeros(H, W, dtype=torch.bool)
    move_mask[10:20, 5:25] = True
    t.train(iterations=2, lr=4e-3, lr_camera=1e-3, lambda_rgb=1.0, lambda_depth=0.1, move_mask=move_mask, densify_interval=500,
            densify_times=0)
    return t, img, depth, move_mask


def test_later_frame_full_stage_matches_the_fit_oracle(reference_trainer, tmp_path, monkeypatch):
    """Frame >= 1, full stage (fit_video.py:288-315): moving Gaussians warped by the flow prior, rgb frozen, xyz of the
    still set frozen, still / flow / scale terms over the right subsets."""
    from gflow_b200 import fit

    ref_trainer, calls = reference_trainer
    W, H = 48, 32
    t, img0, depth0, move_mask = _first_frame(ref_trainer, tmp_path, W, H)
    assert 5 < int((~t.still_mask).sum()) < t.still_mask.numel() - 5, "frame 0 left both still and moving Gaussians"
    g = torch.Generator().manual_seed(3)
    img1 = torch.roll(img0, shifts=1, dims=1).contiguous()
    depth1 = (depth0 * 1.02).contiguous()
    gt_flow = torch.zeros(H, W, 2)
    gt_flow[..., 0] = 1.0 + 0.2 * torch.rand(H, W, generator=g)
    gt_flow[..., 1] = 0.3 * torch.randn(H, W, generator=g)
    t.set_gt_image(img1)
    t.set_gt_depth(depth1)
    t.set_gt_flow(gt_flow)
    state = dict(still_mask=t.still_mask.clone(), last_still_mask=t.last_still_mask.clone(), last_uv=t.last_uv.clone(),
                 last_xyz=t.last_xyz.clone())
    xyz_before = t._attributes["xyz"].detach().clone()
    snap = {}
    orig_add = ref_trainer.SimpleGaussian.add_optimizer

    def add_optimizer(self, *a, **k):  # called right after the pre-update warp (trainer.py:383)
        snap.update({k2: v.detach().clone() for k2, v in self._attributes.items()})
        snap["pose"] = self.pose.detach().clone()
        return orig_add(self, *a, **k)

    monkeypatch.setattr(ref_trainer.SimpleGaussian, "add_optimizer", add_optimizer)
    _Bar.posted = []
    iters = 3
    lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.2, lambda_scale=0.05, lambda_still=0.3, lambda_flow=0.01)
    t.train(iterations=iters, lr=2e-3, lr_camera=0.0, mask=torch.zeros(H, W, 1), move_mask=move_mask, densify_interval=500,
            densify_times=0, **lam)
    # --- our host-side warp == the reference's pre-update processing
    prev = fit.PrevFrame(last_xyz=state["last_xyz"], last_still_mask=state["last_still_mask"], last_uv=state["last_uv"],
                         gt_flow=gt_flow)
    extr0 = FR.pose_to_extr(snap["pose"])
    warped = fit.warp_moving_by_flow(xyz_before, prev, depth1, t.intr, extr0, W, H)
    assert not torch.equal(snap["xyz"], xyz_before), "the reference moved the moving Gaussians"
    assert torch.allclose(warped, snap["xyz"], rtol=1e-4, atol=1e-5)
    # --- the loop itself
    start = {k: snap[k] for k in FR.ATTRS}
    cfg = FR.FitRefConfig(iterations=iters, lr=2e-3, lr_camera=0.0, use_ssim=True, depth_den_min=0.0, freeze_rgb=True, **lam)
    pr = dict(last_xyz=state["last_xyz"], last_still_mask=state["last_still_mask"], last_uv=state["last_uv"], gt_flow=gt_flow,
              and_mask=FR.flow_and_mask(state["last_uv"], W, H, state["still_mask"], False))
    assert int(pr["and_mask"].sum()) > 3
    raw, pose, ab, hist = FR.fit_loop(start, snap["pose"], t.intr, img1, depth1, W, H, cfg, still_mask=state["still_mask"], prev=pr)
    for i in range(iters):
        p, h = _Bar.posted[i], hist[i]
        assert abs(float(p["total"]) - float(h["total"])) <= 1e-4 * abs(float(h["total"])), (i, p, h["total"])
        for key in ("still", "flow", "scale", "var", "depth"):  # posted as "%.6f" strings (trainer.py:464-530)
            assert abs(float(p[key]) - float(h[key])) <= 1.5e-6 + 1e-4 * abs(float(h[key])), (i, key, p[key], h[key])
    for k in FR.ATTRS:
        assert torch.allclose(t._attributes[k].detach(), raw[k], rtol=1e-4, atol=1e-5), k
    assert torch.equal(t._attributes["rgb"].detach(), start["rgb"]), "rgb is frozen on frames >= 1"
    n = state["still_mask"].shape[0]
    assert torch.equal(t._attributes["xyz"].detach()[:n][state["still_mask"]], start["xyz"][:n][state["still_mask"]])
    assert torch.allclose(t.pose.detach(), pose, atol=1e-6)


def test_later_frame_camera_only_stage_matches_the_fit_oracle(reference_trainer, tmp_path, monkeypatch):
    """Frame >= 1, camera-only stage (fit_video.py:256-278): attributes frozen, pose optimised, and every iteration the
    tentatively-moving Gaussians are re-rendered and their footprint leaves the losses (trainer.py:427-455,484)."""
    ref_trainer, calls = reference_trainer
    W, H = 48, 32
    t, img0, depth0, move_mask = _first_frame(ref_trainer, tmp_path, W, H)
    g = torch.Generator().manual_seed(4)
    img1 = torch.roll(img0, shifts=1, dims=1).contiguous()
    depth1 = (depth0 * 1.02).contiguous()
    gt_flow = torch.zeros(H, W, 2)
    gt_flow[..., 0] = 1.0 + 0.2 * torch.rand(H, W, generator=g)
    t.set_gt_image(img1)
    t.set_gt_depth(depth1)
    t.set_gt_flow(gt_flow)
    state = dict(still_mask=t.still_mask.clone(), tentative=t.still_mask_tentative.clone(), last_still_mask=t.last_still_mask.clone(),
                 last_uv=t.last_uv.clone(), last_xyz=t.last_xyz.clone())
    start = {k: v.detach().clone() for k, v in t._attributes.items()}
    pose0 = t.pose.detach().clone()
    move_mask1 = torch.zeros(H, W, dtype=torch.bool)
    move_mask1[2:6, 30:40] = True
    _Bar.posted = []
    calls.clear()
    iters = 3
    lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_flow=0.01)
    t.train(iterations=iters, lr_camera=2e-3, camera_only=True, move_mask=move_mask1, lambda_var=0.0, lambda_still=0.0,
            densify_interval=500, densify_times=0, **lam)
    # two operator chains per iteration: the full set (4 blends) and the moving subset (1 blend)
    chain = ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian"]
    assert calls[: 13 * iters] == (chain + ["alpha_blending"] * 4 + chain + ["alpha_blending"]) * iters
    lr_default = inspect.signature(ref_trainer.SimpleGaussian.train).parameters["lr"].default
    cfg = FR.FitRefConfig(iterations=iters, lr=lr_default, lr_camera=2e-3, use_ssim=True, depth_den_min=0.0, camera_only=True,
                          freeze_rgb=True, lambda_var=0.0, lambda_still=0.0, **lam)
    pr = dict(last_xyz=state["last_xyz"], last_still_mask=state["last_still_mask"], last_uv=state["last_uv"], gt_flow=gt_flow,
              and_mask=FR.flow_and_mask(state["last_uv"], W, H, state["still_mask"], True))
    raw, pose, ab, hist = FR.fit_loop(start, pose0, t.intr, img1, depth1, W, H, cfg, pixel_mask=~move_mask1,
                                      still_mask=state["still_mask"], prev=pr, tentative_still=state["tentative"])
    for i in range(iters):
        p, h = _Bar.posted[i], hist[i]
        assert abs(float(p["total"]) - float(h["total"])) <= 1e-4 * abs(float(h["total"])), (i, p, h["total"])
        assert abs(float(p["depth"]) - float(h["depth"])) <= 1.5e-6 + 1e-4 * float(h["depth"])
        assert abs(float(p["flow"]) - float(h["flow"])) <= 1.5e-6 + 1e-4 * float(h["flow"])
        assert int((~h["pixel_mask"]).sum()) > int(move_mask1.sum()), "the moving footprint widened the mask"
    assert all(torch.equal(t._attributes[k].detach(), start[k]) for k in FR.ATTRS), "attributes are frozen"
    assert not torch.equal(t.pose.detach(), pose0)
    assert torch.allclose(t.pose.detach(), pose, rtol=1e-4, atol=1e-6)
    assert torch.allclose(torch.cat([t.depth_a.detach(), t.depth_b.detach()]), ab, rtol=1e-4, atol=1e-6)


def test_two_frame_sequence_bookkeeping_matches_the_reference(reference_trainer, tmp_path, monkeypatch):
    """gflow_b200.sequence.SequenceFitter (frame loop of fit_video.py:104-349 over the native iteration, here executed
    through the SIMT shim) against the unmodified reference run the same way: frame 0, then camera-only + full stage on
    frame 1.  Compared after every stage: the discrete bookkeeping (still / tentative masks), last_uv / last_xyz, the
    first loss of each stage (it sees the flow warp, the masks and the previous-frame state) and the parameters.
    The two runs round differently and Adam's early steps are sign-like, so attribute comparisons allow a small fraction
    of Gaussians with near-zero gradients to differ by a step."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
    import emu
    from gflow_b200 import fit, sequence

    ref_trainer, calls = reference_trainer
    W, H, N = 48, 32, 300
    img0, depth0 = _scene(W, H)
    np.random.seed(0)
    torch.manual_seed(0)
    t = ref_trainer.SimpleGaussian(gt_image=img0, gt_depth=depth0, num_points=N, sequence_path=str(tmp_path / "seq"))
    t.load_camera(focal=0.6 * W, pp=[W / 2.0, H / 2.0], show=False)
    t.init_gaussians_from_image(gt_image=img0, gt_depth=depth0, num_points=N)
    raw0 = {k: v.detach().clone() for k, v in t._attributes.items()}
    pose0 = t.pose.detach().clone()
    mm0 = torch.zeros(H, W, dtype=torch.bool)
    mm0[10:20, 5:25] = True
    mm1 = torch.zeros(H, W, dtype=torch.bool)
    mm1[9:19, 7:27] = True
    g = torch.Generator().manual_seed(5)
    img1, depth1 = torch.roll(img0, shifts=1, dims=1).contiguous(), (depth0 * 1.02).contiguous()
    gt_flow = torch.zeros(H, W, 2)
    gt_flow[..., 0] = 1.0 + 0.2 * torch.rand(H, W, generator=g)
    lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.2, lambda_scale=0.05)
    it0, itc, ita = 3, 2, 3
    # ---------------- ours (native iteration through the SIMT shim)
    monkeypatch.setattr(fit, "NativeFitLoop", emu.fit_loop_class())
    monkeypatch.setattr(fit.FrameFitter, "render", lambda self, bg=0.0, want_depth=True, with_depth=False: (None, None, None))
    cfg = sequence.SequenceConfig(num_points=N, lr=4e-3, lr_camera=1e-3, iterations_first=it0, lr_after=2e-3, iterations_after=ita,
                                  camera_first=True, lr_camera_after=2e-3, iterations_camera=itc, densify_interval=0,
                                  densify_times=0, densify_interval_after=0, densify_times_after=0, lambda_still=0.3,
                                  lambda_flow=0.01, native=True, **lam)
    seq = sequence.SequenceFitter(raw0, t.intr, pose0, W, H, cfg)

    def close(a, b, atol, frac=0.03):
        bad = ((a - b).abs() > atol).any(dim=-1) if a.dim() > 1 else (a - b).abs() > atol
        return float(bad.float().mean()) <= frac

    def compare(stage):
        assert torch.equal(seq.still_mask, t.still_mask), stage
        assert torch.equal(seq.still_mask_tentative, t.still_mask_tentative), stage
        assert close(seq.last_uv, t.last_uv, 0.05), stage
        assert close(seq.last_xyz, t.last_xyz, 1e-3), stage
        for k in FR.ATTRS:
            assert close(seq.attrs[k], t._attributes[k].detach(), 2e-3), (stage, k)
        assert torch.allclose(seq.pose, t.pose.detach(), atol=5e-3), stage

    # ---------------- frame 0
    _Bar.posted = []
    t.train(iterations=it0, lr=4e-3, lr_camera=1e-3, move_mask=mm0, densify_interval=500, densify_times=0, **lam)
    ref_first = [float(p["total"]) for p in _Bar.posted]
    out0 = seq.fit_first(img0, depth0, mm0)
    assert abs(out0.losses["first"][0] - ref_first[0]) <= 1e-4 * ref_first[0]
    assert abs(out0.losses["first"][-1] - ref_first[-1]) <= 2e-2 * ref_first[-1]
    compare("frame 0")
    # ---------------- frame 1: camera-only, then everything
    t.set_gt_image(img1)
    t.set_gt_depth(depth1)
    t.set_gt_flow(gt_flow)
    _Bar.posted = []
    t.train(iterations=itc, lr_camera=2e-3, camera_only=True, move_mask=mm1, lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.0,
            lambda_still=0.0, lambda_flow=0.01, densify_interval=500, densify_times=0)
    ref_cam = [float(p["total"]) for p in _Bar.posted]
    _Bar.posted = []
    t.train(iterations=ita, lr=2e-3, lr_camera=0.0, mask=torch.zeros(H, W, 1), move_mask=mm1, lambda_still=0.3, lambda_flow=0.01,
            densify_interval=500, densify_times=0, **lam)
    ref_all = [float(p["total"]) for p in _Bar.posted]
    out1 = seq.fit_next(img1, depth1, gt_flow, mm1, occ_mask=torch.zeros(H, W, 1))
    assert abs(out1.losses["camera"][0] - ref_cam[0]) <= 2e-2 * ref_cam[0], (out1.losses["camera"], ref_cam)
    assert abs(out1.losses["all"][0] - ref_all[0]) <= 2e-2 * ref_all[0], (out1.losses["all"], ref_all)
    assert abs(out1.losses["all"][-1] - ref_all[-1]) <= 3e-2 * ref_all[-1]
    compare("frame 1")
    assert seq.state().num_points == N


def test_patched_trainer_runs_its_stages_natively_and_keeps_the_reference_state(reference_trainer, tmp_path, monkeypatch):
    """gflow_b200.accelerate.patch(trainer): SimpleGaussian.train keeps its signature, side effects and return tuple but
    the iterations run in csrc/fit.cu (here through the SIMT shim).  Two trainers from the same initial state, one
    patched, are driven the way fit_video.py drives them for two frames (first stage; camera-only + full stage)."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
    import emu
    from gflow_b200 import accelerate, fit

    ref_trainer, calls = reference_trainer
    W, H, N = 48, 32, 300
    monkeypatch.setattr(fit, "NativeFitLoop", emu.fit_loop_class())
    monkeypatch.setattr(fit.FrameFitter, "render", lambda self, bg=0.0, want_depth=True, with_depth=False: (None, None, None))
    t_ref, img0, depth0 = ref_harness.new_trainer(ref_trainer, tmp_path, W, H, N)
    t_nat, _, _ = ref_harness.new_trainer(ref_trainer, tmp_path, W, H, N)  # same seeds -> same initial state
    assert all(torch.equal(t_ref._attributes[k], t_nat._attributes[k]) for k in FR.ATTRS)
    native_train = accelerate.native_train
    mm0 = torch.zeros(H, W, dtype=torch.bool)
    mm0[10:20, 5:25] = True
    mm1 = torch.zeros(H, W, dtype=torch.bool)
    mm1[9:19, 7:27] = True
    g = torch.Generator().manual_seed(5)
    img1, depth1 = torch.roll(img0, shifts=1, dims=1).contiguous(), (depth0 * 1.02).contiguous()
    gt_flow = torch.zeros(H, W, 2)
    gt_flow[..., 0] = 1.0 + 0.2 * torch.rand(H, W, generator=g)
    lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.2, lambda_scale=0.05)

    def both(**kw):
        out_ref = t_ref.train(**kw)
        out_nat = native_train(t_nat, **kw)
        assert len(out_nat) == len(out_ref) == 8
        assert out_nat[0][0].shape == (H, W, 3) and out_nat[0][0].dtype == np.uint8
        return out_ref, out_nat

    def compare(stage, check_masks=True):
        for k in FR.ATTRS:
            a, b = t_nat._attributes[k].detach(), t_ref._attributes[k].detach()
            assert a.shape == b.shape and float(((a - b).abs() > 1e-3).any(dim=1).float().mean()) <= 0.02, (stage, k)
        assert torch.allclose(t_nat.pose.detach(), t_ref.pose.detach(), atol=1e-4), stage
        assert torch.allclose(t_nat.depth_a.detach(), t_ref.depth_a.detach(), atol=1e-4), stage
        if check_masks:
            assert torch.equal(t_nat.still_mask, t_ref.still_mask) and torch.equal(t_nat.still_mask_tentative, t_ref.still_mask_tentative)
            assert torch.allclose(t_nat.last_uv, t_ref.last_uv, atol=2e-2) and t_nat.last_num == t_ref.last_num
            assert torch.allclose(t_nat.last_xyz, t_ref.last_xyz, atol=1e-3)

    both(iterations=3, lr=4e-3, lr_camera=1e-3, move_mask=mm0, densify_interval=500, densify_times=0, save_ckpt=True, ckpt_name="0000",
         **lam)
    compare("frame 0")
    assert os.path.exists(os.path.join(t_nat.dir, "ckpt", "0000.tar")), "the reference's own save_checkpoint ran"
    for t in (t_ref, t_nat):
        t.set_gt_image(img1)
        t.set_gt_depth(depth1)
        t.set_gt_flow(gt_flow)
    both(iterations=2, lr_camera=2e-3, camera_only=True, move_mask=mm1, lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.0, lambda_still=0.0,
         lambda_flow=0.01, densify_interval=500, densify_times=0)
    compare("frame 1 camera")
    out_ref, out_nat = both(iterations=3, lr=2e-3, lr_camera=0.0, mask=torch.zeros(H, W, 1), move_mask=mm1, lambda_still=0.3,
                            lambda_flow=0.01, densify_interval=500, densify_times=0, **lam)
    compare("frame 1 all")
    assert out_nat[3] is not None and out_nat[5] is not None  # still / moving renders exist once a still mask exists
    # the patch itself
    accelerate.patch(ref_trainer)
    assert ref_trainer.SimpleGaussian.train is native_train and hasattr(ref_trainer.SimpleGaussian, "train_reference")
    ref_trainer.SimpleGaussian.train = ref_trainer.SimpleGaussian.train_reference
    del ref_trainer.SimpleGaussian.train_reference, ref_trainer.SimpleGaussian._gflow_b200_native
