#!/usr/bin/env python
"""In-situ per-kernel durations of the native fit iteration at BASELINE config 3 (torch.profiler / CUPTI):
python tools/fit_kernel_times.py [iterations] [ssim]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gflow_b200 import fit  # noqa: E402
from gflow_b200.synthetic import make_scene  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ssim = len(sys.argv) > 2 and sys.argv[2] == "ssim"
N, W, H = 60000, 854, 480
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0, profile="gflow")
raw = {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
       "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
       "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}
raw = {k: v.to(dev) for k, v in raw.items()}
f = fit.FrameFitter(raw, sc.intr.to(dev), fit.extr_to_pose(sc.extr).to(dev), W, H)
with torch.no_grad():
    img, dmap, _ = f.render(0.0, want_depth=True)
gt_image, gt_depth = img.permute(1, 2, 0).contiguous() * 0.9, dmap.permute(1, 2, 0).contiguous()
cfg = fit.FitConfig(iterations=iters + 10, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, use_ssim=ssim, native=True)
loop = fit.NativeFitLoop(f, gt_image, gt_depth, cfg)
loop.run(10)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    loop.run(iters)
    torch.cuda.synchronize()
tot = collections.OrderedDict()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        d = tot.setdefault(name, [0.0, 0])
        d[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        d[1] += 1
print(f"# native fit iteration, {N} Gaussians {W}x{H}, ssim={ssim}, avg per iteration over {iters}")
s = 0.0
for name, (t, n) in tot.items():
    print(f"{t / iters:9.2f} us/iter  x{n / iters:4.1f}  {name}")
    s += t / iters
print(f"{s:9.2f} us/iter  total GPU busy")
