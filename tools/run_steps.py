#!/usr/bin/env python
"""Run a few render steps (for ncu captures): python tools/run_steps.py [fused|chain] [steps] [workload] [profile]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gflow_b200 as G  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "fused"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, W, H = CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "cfg2"]
profile = sys.argv[4] if len(sys.argv) > 4 else "synthetic"
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0, profile=profile)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
fn = G.rasterization if path == "fused" else G.rasterization_unfused
for _ in range(steps):
    for p in ps:
        p.grad = None
    img = fn(*ps, intr, extr, W, H, 0.0)
    (img * Gimg).sum().backward()
torch.cuda.synchronize()
print("done", path, steps)
