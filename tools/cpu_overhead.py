#!/usr/bin/env python
"""Host-side cost of one fused render step: wall time of the forward call (returns once K has landed),
of the backward call, and of the python in between (GPU idle at the start of every step)."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gflow_b200 as G  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

N, W, H = CONFIGS["cfg2"]
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
for path, fn in (("fused", G.rasterization), ("chain", G.rasterization_unfused)):
    tf, tb, tt = [], [], []
    for it in range(220):
        for p in ps:
            p.grad = None
        extr.grad = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        img = fn(*ps, intr, extr, W, H, 0.0)
        t1 = time.perf_counter()
        img.backward(Gimg)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it >= 20:
            tf.append(t1 - t0), tb.append(t2 - t1), tt.append(t3 - t0)
    us = lambda v: round(1e6 * statistics.median(v), 1)  # noqa: E731
    print(f"{path}: forward call {us(tf)} us, backward call {us(tb)} us, step incl. final sync {us(tt)} us (medians, GPU idle at step start)")
