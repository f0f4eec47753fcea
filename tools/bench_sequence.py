#!/usr/bin/env python
"""A short synthetic video fitted the way scripts/fit_video.sh does (500 iterations on frame 0 with two
densifications, then 150 camera-only + 300 full iterations per frame), operator path or native kernels:

  python tools/bench_sequence.py [--frames 3] [--points 50000] [--operator] [--scale 1.0]

Prints one JSON line: total iterations, seconds, iterations/s, Gaussians per frame."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gflow_b200 import fit, sequence  # noqa: E402
from gflow_b200.synthetic import make_camera, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--points", type=int, default=50000)
ap.add_argument("--operator", action="store_true", help="msplat operators + autograd + torch.optim (default: native kernels)")
ap.add_argument("--scale", type=float, default=1.0, help="scales every iteration count (quick runs)")
ap.add_argument("--size", type=int, nargs=2, default=[854, 480], metavar=("W", "H"))
args = ap.parse_args()
dev = torch.device("cuda:0")
W, H = args.size
sc = make_scene(args.points, W, H, seed=0, profile="gflow")
raw = {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
       "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
       "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}
raw = {k: v.to(dev) for k, v in raw.items()}
intr = sc.intr.to(dev)


def target(i):
    """The scene seen from a slowly drifting camera: image, depth map, a constant flow, a box of 'moving' pixels."""
    gen = torch.Generator().manual_seed(100)
    _, extr = make_camera(W, H, gen)
    extr = extr.clone()
    extr[0, 3] += 0.01 * i
    f = fit.FrameFitter(raw, intr, fit.extr_to_pose(extr).to(dev), W, H)
    with torch.no_grad():
        img, dmap, _ = f.render(0.0, want_depth=True)
    move = torch.zeros(H, W, dtype=torch.bool, device=dev)
    move[H // 3:2 * H // 3, W // 4 + 5 * i:W // 2 + 5 * i] = True
    flow = torch.zeros(H, W, 2, device=dev)
    flow[..., 0] = 2.5
    return img.permute(1, 2, 0).contiguous(), dmap.permute(1, 2, 0).contiguous().clamp_min(0.05), flow, move


s = args.scale
cfg = sequence.SequenceConfig(num_points=args.points, native=not args.operator, iterations_first=max(2, int(500 * s)),
                              iterations_camera=max(2, int(150 * s)), iterations_after=max(2, int(300 * s)),
                              densify_interval=max(1, int(150 * s)), densify_interval_after=max(1, int(100 * s)))
gen = torch.Generator().manual_seed(7)
start = dict(raw, rgb=torch.zeros_like(raw["rgb"]))  # grey start: there is something to fit
_, extr0 = make_camera(W, H, torch.Generator().manual_seed(100))
seq = sequence.SequenceFitter(start, intr, fit.extr_to_pose(extr0).to(dev), W, H, cfg)
targets = [target(i) for i in range(args.frames)]
torch.cuda.synchronize()
t0 = time.perf_counter()
counts, first, last = [], None, None
img, dep, flow, move = targets[0]
out = seq.fit_first(img, dep, move)
first = out.losses["first"][0]
counts.append(out.num_points)
iters = cfg.iterations_first
for i in range(1, args.frames):
    img, dep, flow, move = targets[i]
    out = seq.fit_next(img, dep, flow, move, occ_mask=move.float().unsqueeze(-1))
    counts.append(out.num_points)
    iters += cfg.iterations_camera + cfg.iterations_after
    last = out.losses["all"][-1]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"metric": "fit_video-style sequence, iterations / s", "value": iters / dt, "unit": "iters/s", "frames": args.frames,
                  "iterations": iters, "seconds": dt, "native": not args.operator, "gaussians_per_frame": counts,
                  "loss_first": first, "loss_last": last, "resolution": [W, H]}))
