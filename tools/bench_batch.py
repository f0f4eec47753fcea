#!/usr/bin/env python
"""Frames/s of gflow_b200.BatchedRenderStep against one GraphedRenderStep after the other:
python tools/bench_batch.py [workload] [frames] [profile]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gflow_b200 as G
from gflow_b200.synthetic import CONFIGS, make_camera, make_grad_image, make_scene

N, W, H = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
profile = sys.argv[3] if len(sys.argv) > 3 else "synthetic"
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0, profile=profile)
ps = [t.to(dev) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
cams = [make_camera(W, H, torch.Generator().manual_seed(2000 + f)) for f in range(F)]
Gimg = make_grad_image(3, W, H).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


one = G.GraphedRenderStep(*ps, cams[0][0].to(dev), cams[0][1].to(dev), W, H, sc.bg)
one.g_image.copy_(Gimg)
ms_one = timed(one)
batch = G.BatchedRenderStep(*ps, torch.stack([c[0] for c in cams]).to(dev), torch.stack([c[1] for c in cams]).to(dev), W, H, sc.bg)
for g in batch.g_images:
    g.copy_(Gimg)
ms_batch = timed(batch)
batch.check()
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'cfg2'} {profile}: one frame {ms_one * 1e3:.1f} us = {1e3 / ms_one:.0f} frames/s;  "
      f"batch of {F}: {ms_batch * 1e3:.1f} us = {F * 1e3 / ms_batch:.0f} frames/s  ({F * ms_one / ms_batch:.2f} x)")
