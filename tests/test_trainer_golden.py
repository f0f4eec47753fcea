"""tests/golden/trainer_stages.npz holds what the UNMODIFIED reference loop (/root/reference/gflow/trainer.py, run by
tests/golden/make_trainer_golden.py) rendered from and posted, iteration by iteration, for a first-frame stage, a
later-frame camera-only stage and a later-frame full stage.  Here, without /root/reference:
  * oracle/fit_ref.py reproduces every posted loss from the recorded state of that iteration,
  * the native loop (csrc/fit.cu through the SIMT shim) reproduces each stage from its recorded start,
  * the host-side flow warp reproduces the reference's pre-update step.
The same comparison runs on the GPU in tests/test_zz_native_fit_gpu.py."""
import os
import sys

import pytest
import torch

import fit_check
from gflow_b200 import fit
from oracle import fit_ref as FR

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import emu  # noqa: E402

STAGES = ["first", "camera", "all"]


@pytest.mark.parametrize("stage", STAGES)
def test_fit_oracle_reproduces_every_posted_loss(stage):
    G = fit_check.load_trainer_golden()
    W, H = int(G["W"]), int(G["H"])
    cfg, _, _, img, depth, kw = fit_check.golden_stage_inputs(G, stage)
    rcfg = fit_check.ref_config(cfg)
    rcfg.depth_den_min = 0.0  # the reference does not clamp the depth-loss denominator
    prev = kw["prev"]
    if prev is not None:
        prev = dict(prev, and_mask=FR.flow_and_mask(prev["last_uv"], W, H, kw["still_mask"], cfg.camera_only))
    keep = kw["pixel_mask"]
    for i in range(cfg.iterations):
        raw = {k: G[f"{stage}/state/{k}"][i] for k in FR.ATTRS}
        pose, ab = G[f"{stage}/state/pose"][i], G[f"{stage}/state/ab"][i]
        if stage == "camera":  # cumulative footprint of the tentatively-moving Gaussians (trainer.py:427-451)
            keep = keep & ~FR.moving_footprint(raw, pose, G["intr"], W, H, 0.0, kw["tentative_still"])
        loss, parts = FR.iteration_loss(raw, pose, ab, G["intr"], img, depth, keep, W, H, rcfg, prev, kw["still_mask"])
        ref = float(G[f"{stage}/posted/total"][i])
        assert abs(float(loss) - ref) <= 1e-5 * abs(ref), (stage, i, float(loss), ref)
        for key, ours in (("depth", "depth"), ("var", "var"), ("scale", "scale"), ("still", "still"), ("flow", "flow")):
            posted = float(G[f"{stage}/posted/{key}"][i])
            if posted == posted and ours in parts:  # posted as "%.6f" strings
                assert abs(float(parts[ours]) - posted) <= 1.5e-6, (stage, i, key)


@pytest.mark.parametrize("stage", STAGES)
def test_native_loop_reproduces_the_recorded_stage(stage):
    # the emulated kernels follow the reference's trajectory to ~1e-6; the GPU run uses the function's looser defaults
    fit_check.check_native_stage_against_reference_golden(emu.fit_loop_class(), "cpu", stage, loss_rtol=1e-4, attr_atol=1e-3,
                                                          attr_frac=0.0, pose_atol=1e-5)


def test_flow_warp_reproduces_the_reference_pre_update():
    G = fit_check.load_trainer_golden()
    W, H = int(G["W"]), int(G["H"])
    prev = fit.PrevFrame(last_xyz=G["all/last_xyz"], last_still_mask=G["all/last_still_mask"].bool(), last_uv=G["all/last_uv"],
                         gt_flow=G["gt_flow"])
    before = G["camera/final/xyz"]  # the camera-only stage leaves the attributes alone
    extr = fit.pose_to_extr(G["all/state/pose"][0])
    warped = fit.warp_moving_by_flow(before, prev, G["depth1"], G["intr"], extr, W, H)
    after = G["all/state/xyz"][0]
    assert not torch.equal(before, after)
    assert torch.allclose(warped, after, rtol=1e-4, atol=1e-5)


def test_operator_path_reproduces_the_recorded_stages():
    """The operator path (FrameFitter.train, native=False: operators + autograd + torch.optim.Adam, the way
    gflow/trainer.py:387-582 drives the drop-in) against the same recording, with the operators answered by the CPU
    oracle (tests/simt/fake_cuda_run.py); the GPU suite runs the identical check body on the real kernels."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, os.path.join(here, "simt", "fake_cuda_run.py"),
                          os.path.join(here, "simt", "operator_golden_check.py")], capture_output=True, text=True, cwd=here,
                         timeout=900)
    assert res.returncode == 0 and "OPERATOR_GOLDEN_OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]
