"""Runs the native-fit GPU checks in a process of its own and prints one JSON object {case: "ok" | message}.

tests/test_zz_native_fit_gpu.py launches this file: a CUDA fault in a kernel that has not yet run on
hardware would otherwise poison the CUDA context of the whole pytest process."""
import json
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch  # noqa: E402

import fit_check  # noqa: E402
from gflow_b200 import fit  # noqa: E402
from gflow_b200.synthetic import make_scene  # noqa: E402


def native_vs_operator_path():
    """FrameFitter.train through both execution modes from the same start: same loss curve to a few percent
    (the two paths round differently and Adam's first steps are sign-like, so this is a trajectory check;
    per-iteration parity is what fit_check.run_and_check asserts)."""
    dev = torch.device("cuda:0")
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N=6000, W=320, H=200, seed=11)
    out = {}
    for native in (False, True):
        f = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), pose.to(dev), 320, 200)
        cfg = fit.FitConfig(iterations=40, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, native=native,
                            check_every=16)
        res = f.train(gt_image.to(dev), gt_depth.to(dev), cfg)
        out[native] = res
        assert len(res.losses) == 40 and all(v == v for v in res.losses)
    a, b = out[False].losses, out[True].losses
    assert abs(a[0] - b[0]) <= 1e-3 * abs(a[0]), (a[0], b[0])
    assert b[-1] < 0.9 * b[0], "native loop reduces the loss"
    assert abs(a[-1] - b[-1]) <= 0.05 * abs(a[-1]), (a[-1], b[-1])


def densification_both_paths():
    """trainer.py:566-571: (iteration + 1) % interval == 0 appends Gaussians drawn from the error map; both
    execution modes grow the attribute tensors and keep optimising with finite losses."""
    dev = torch.device("cuda:0")
    sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N=3000, W=160, H=120, seed=21)
    for native in (False, True):
        f = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), pose.to(dev), 160, 120)
        cfg = fit.FitConfig(iterations=13, lr=4e-3, lambda_depth=0.1, native=native, densify_interval=4, densify_times=2,
                            densify_err_thre=1e-5, densify_err_percent=0.3, check_every=3)
        res = f.train(gt_image.to(dev), gt_depth.to(dev), cfg)
        n = f.attrs["xyz"].shape[0]
        assert n > 3000 and all(f.attrs[k].shape[0] == n for k in fit.ATTRS), (native, n)
        assert len(res.losses) == 13 and all(v == v for v in res.losses), (native, res.losses)
        assert res.image.shape == (3, 120, 160)


def concurrent_frames_on_streams():
    """Two frames side by side on two CUDA streams follow the same loss curves as one after the other (float
    atomics reorder sums, so the comparison is a tolerance, not bit equality)."""
    dev = torch.device("cuda:0")
    probs = [fit_check.make_problem(N=4000, W=256, H=160, seed=40 + i) for i in range(2)]
    cfg = fit.FitConfig(iterations=30, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, native=True, check_every=8)

    def fitters():
        return [fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), pose.to(dev), 256, 160)
                for sc, raw, pose, _, _ in probs]

    alone = []
    for f, (_, _, _, gi, gd) in zip(fitters(), probs):
        lp = fit.NativeFitLoop(f, gi.to(dev), gd.to(dev), cfg)
        lp.run(30)
        alone.append(lp.loss_history()[:, 0].cpu())
    loops = fit.fit_frames_concurrently(fitters(), [(p[3].to(dev), p[4].to(dev)) for p in probs], cfg)
    torch.cuda.synchronize()
    for a, lp in zip(alone, loops):
        b = lp.loss_history()[:, 0].cpu()
        assert torch.allclose(a, b, rtol=2e-2), (a, b)
        assert abs(float(a[0]) - float(b[0])) <= 1e-4 * float(a[0])


def native_iterations_at_config3_size():
    """Two native iterations at BASELINE config 3's size (60 000 Gaussians, 854x480, mse + 1-SSIM + depth, camera
    learning on) checked one by one against oracle/fit_ref.py: loss terms, raw-attribute gradients, dL/dpose and the
    in-kernel Adam update (fit_check.run_and_check).  The oracle needs ~12 s per iteration on the host."""
    from gflow_b200.fit import NativeFitLoop

    cfg = fit.FitConfig(iterations=300, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, native=True)
    fit_check.run_and_check(NativeFitLoop, "cuda:0", cfg, n_iters=2, N=60000, W=854, H=480, seed=0, capacity=400000)


def main():
    results = {}
    cases = fit_check.case_list()
    from gflow_b200.fit import NativeFitLoop

    for name, cfg, kwargs in cases:
        try:
            loop, fitter, raw0, pose0 = fit_check.run_and_check(NativeFitLoop, "cuda:0", cfg, **kwargs)
            fit_check.post_checks(name, loop, fitter, raw0, pose0, kwargs)
            torch.cuda.synchronize()
            results[name] = "ok"
        except Exception:  # noqa: BLE001
            results[name] = traceback.format_exc()[-1500:]
    for stage in ("first", "camera", "all"):  # what the unmodified reference loop produced (tests/golden/trainer_stages.npz)
        try:
            fit_check.check_native_stage_against_reference_golden(NativeFitLoop, "cuda:0", stage)
            torch.cuda.synchronize()
            results[f"reference_golden_{stage}"] = "ok"
        except Exception:  # noqa: BLE001
            results[f"reference_golden_{stage}"] = traceback.format_exc()[-1500:]
    for stage in ("first", "camera", "all"):  # the same recorded stages through the operator path
        try:
            fit_check.check_operator_stage_against_reference_golden("cuda:0", stage)
            torch.cuda.synchronize()
            results[f"operator_golden_{stage}"] = "ok"
        except Exception:  # noqa: BLE001
            results[f"operator_golden_{stage}"] = traceback.format_exc()[-1500:]
    try:
        native_iterations_at_config3_size()
        torch.cuda.synchronize()
        results["native_config3_size"] = "ok"
    except Exception:  # noqa: BLE001
        results["native_config3_size"] = traceback.format_exc()[-1500:]
    try:
        native_vs_operator_path()
        torch.cuda.synchronize()
        results["native_vs_operator_path"] = "ok"
    except Exception:  # noqa: BLE001
        results["native_vs_operator_path"] = traceback.format_exc()[-1500:]
    try:
        densification_both_paths()
        torch.cuda.synchronize()
        results["densification_both_paths"] = "ok"
    except Exception:  # noqa: BLE001
        results["densification_both_paths"] = traceback.format_exc()[-1500:]
    try:
        concurrent_frames_on_streams()
        results["concurrent_frames_on_streams"] = "ok"
    except Exception:  # noqa: BLE001
        results["concurrent_frames_on_streams"] = traceback.format_exc()[-1500:]
    print("RESULT " + json.dumps(results), flush=True)


if __name__ == "__main__":
    main()
