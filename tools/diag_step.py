#!/usr/bin/env python
"""Host time vs GPU time of the render step: python tools/diag_step.py [steps] [flush 0|1]"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gflow_b200 as G
from gflow_b200.synthetic import make_grad_image, make_scene
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
flush = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
dev = torch.device("cuda:0")
N, W, H = 60000, 854, 480
sc = make_scene(N, W, H, seed=0)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    for p in ps: p.grad = None
    extr.grad = None
    img = G.rasterization(*ps, intr, extr, W, H, 0.0)
    img.backward(Gimg)
for _ in range(10): step()
torch.cuda.synchronize()
for blk in range(6):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    host = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, b in ev:
        if flush: buf.fill_(1)
        a.record()
        h0 = time.perf_counter()
        step()
        host.append(time.perf_counter() - h0)
        b.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    ms = [a.elapsed_time(b) for a, b in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(steps - 1)]
    print(f"block {blk}: gpu step median {statistics.median(ms)*1e3:.1f} us mean {statistics.mean(ms)*1e3:.1f}  gap(between steps) median {statistics.median(gaps)*1e3:.1f} us | host per step() median {statistics.median(host)*1e6:.1f} us mean {statistics.mean(host)*1e6:.1f} max {max(host)*1e6:.0f} | enqueue wall {t_enq/steps*1e6:.1f} us/step, total wall {t_all/steps*1e6:.1f} us/step  mem {torch.cuda.memory_allocated()>>20} MiB reserved {torch.cuda.memory_reserved()>>20} MiB")
