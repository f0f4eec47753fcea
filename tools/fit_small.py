#!/usr/bin/env python
"""A few iterations of the native fit loop (csrc/fit.cu) with every loss term on, for compute-sanitizer / ncu:
python tools/fit_small.py [small|cfg2] [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gflow_b200 import fit  # noqa: E402
from gflow_b200.synthetic import make_scene  # noqa: E402

size = sys.argv[1] if len(sys.argv) > 1 else "small"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, W, H = (60000, 854, 480) if size == "cfg2" else ((301, 64, 40) if size == "tiny" else (3001, 200, 120))  # odd N: unaligned tails
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0, profile="gflow" if size == "cfg2" else "synthetic")
raw = {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
       "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
       "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}
raw = {k: v.to(dev) for k, v in raw.items()}
pose = fit.extr_to_pose(sc.extr).to(dev)
f = fit.FrameFitter(raw, sc.intr.to(dev), pose, W, H)
with torch.no_grad():
    img, dmap, uv = f.render(0.0, want_depth=True)
gt_image, gt_depth = img.permute(1, 2, 0).contiguous() * 0.9, dmap.permute(1, 2, 0).contiguous() * 1.1
g = torch.Generator().manual_seed(1)
prev = fit.PrevFrame(last_xyz=raw["xyz"] + 0.01, last_still_mask=(torch.rand(N, generator=g) > 0.5).to(dev), last_uv=uv.detach(),
                     gt_flow=torch.randn(H, W, 2, generator=g).to(dev))
cfg = fit.FitConfig(iterations=max(iters, 2), lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=True, lambda_var=0.1,
                    lambda_scale=0.1, lambda_still=0.1, lambda_flow=0.01, native=True, check_every=2)
res = f.train(gt_image, gt_depth, cfg, pixel_mask=(torch.rand(H, W, generator=g) > 0.1).to(dev),
              still_mask=(torch.rand(N - 7, generator=g) > 0.5).to(dev), prev=prev)
torch.cuda.synchronize()
print("losses", [round(v, 6) for v in res.losses])
assert all(v == v for v in res.losses)
if size in ("small", "tiny"):
    # the two other stage kinds: densification between iterations, and a camera-only stage whose moving subset is
    # re-rendered every iteration (second pipeline pass + fit_move_mask)
    cfg2 = fit.FitConfig(iterations=6, lr=4e-3, lambda_depth=0.1, native=True, densify_interval=2, densify_times=2,
                         densify_err_thre=1e-6, densify_err_percent=0.5, check_every=2)
    f2 = fit.FrameFitter(raw, sc.intr.to(dev), pose, W, H)
    r2 = f2.train(gt_image, gt_depth, cfg2, prev=prev, occlusion_mask=(torch.rand(H, W, 1, generator=g) > 0.7).float().to(dev),
                  still_mask=(torch.rand(N - 7, generator=g) > 0.5).to(dev))
    torch.cuda.synchronize()
    print("densify: N", N, "->", f2.attrs["xyz"].shape[0], "losses", [round(v, 6) for v in r2.losses])
    assert f2.attrs["xyz"].shape[0] > N and all(v == v for v in r2.losses)
    cfg3 = fit.FitConfig(iterations=4, lr_camera=1e-3, lambda_depth=0.1, lambda_flow=0.01, use_ssim=True, camera_only=True,
                         native=True, check_every=2)
    f3 = fit.FrameFitter(raw, sc.intr.to(dev), pose, W, H)
    r3 = f3.train(gt_image, gt_depth, cfg3, prev=prev, pixel_mask=(torch.rand(H, W, generator=g) > 0.1).to(dev),
                  still_mask=(torch.rand(N - 7, generator=g) > 0.5).to(dev),
                  tentative_still=(torch.rand(N - 7, generator=g) > 0.2).to(dev))
    torch.cuda.synchronize()
    print("camera-only losses", [round(v, 6) for v in r3.losses])
    assert all(v == v for v in r3.losses)
