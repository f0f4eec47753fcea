#!/usr/bin/env python
"""The operator chain (render.py:21-64 pattern) alone: python tools/diag_chain.py [steps]   (under ncu: a launch list)"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gflow_b200 as G
from gflow_b200 import capi
from gflow_b200.synthetic import make_grad_image, make_scene
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
N, W, H = 60000, 854, 480
sc = make_scene(N, W, H, seed=0)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
lib = capi.load()
def step():
    for p in ps: p.grad = None
    extr.grad = None
    img = G.rasterization_unfused(*ps, intr, extr, W, H, 0.0)
    img.backward(Gimg)
for _ in range(5): step()
torch.cuda.synchronize()
l0 = lib.gfb_kernel_launch_count()
step()
torch.cuda.synchronize()
print("library launches per chain step:", lib.gfb_kernel_launch_count() - l0)
for blk in range(3):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, b in ev:
        a.record(); step(); b.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    ms = [a.elapsed_time(b) for a, b in ev]
    print(f"block {blk}: host enqueue {1e6*t_enq/steps:.1f} us/step, wall {1e6*t_all/steps:.1f} us/step, event median {1e3*statistics.median(ms):.1f} us")
