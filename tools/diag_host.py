#!/usr/bin/env python
"""Where the host time of one render step goes (fused op and operator chain): python tools/diag_host.py"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gflow_b200 as G
from gflow_b200 import capi
from gflow_b200.synthetic import make_grad_image, make_scene
dev = torch.device("cuda:0")
N, W, H = 60000, 854, 480
sc = make_scene(N, W, H, seed=0)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
lib = capi.load()

def measure(label, raster, reps=200, mt=True):
    torch.autograd.set_multithreading_enabled(mt)
    tf, tb, tz = [], [], []
    for i in range(reps + 20):
        t0 = time.perf_counter()
        for p in ps: p.grad = None
        extr.grad = None
        t1 = time.perf_counter()
        img = raster(*ps, intr, extr, W, H, 0.0)
        t2 = time.perf_counter()
        img.backward(Gimg)
        t3 = time.perf_counter()
        if i % 8 == 7: torch.cuda.synchronize()   # keep the launch queue shallow: pure host cost
        if i >= 20:
            tz.append(t1 - t0); tf.append(t2 - t1); tb.append(t3 - t2)
    torch.cuda.synchronize()
    m = lambda v: statistics.median(v) * 1e6
    print(f"{label:38s} zero grads {m(tz):5.1f} us | forward call {m(tf):6.1f} us | backward call {m(tb):6.1f} us | total {m(tz)+m(tf)+m(tb):6.1f} us")

measure("fused, engine multithreading on", G.rasterization)
measure("fused, engine multithreading off", G.rasterization, mt=False)
measure("chain, engine multithreading on", G.rasterization_unfused)
measure("chain, engine multithreading off", G.rasterization_unfused, mt=False)
torch.autograd.set_multithreading_enabled(True)
# raw C-ABI launches without torch: forward + backward through ops internals is not separable here; time a trivial custom op
class Nop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *a):
        ctx.n = len(a)
        return a[0].new_empty(3, H, W)
    @staticmethod
    def backward(ctx, g):
        return tuple(torch.empty_like(p) for p in ps) + (None, torch.empty_like(extr))
def nop(*a): return Nop.apply(*a[:7])
measure("python no-op Function (autograd floor)", lambda *a: nop(*a))

# ---- per-operator host time of the chain's forward (queue kept shallow)
def per_op(reps=200):
    names = ["project_point", "depth != 0", "compute_cov3d", "ewa_project", "sort_gaussian", "alpha_blending"]
    acc = {n: [] for n in names}
    xyz, scale, rot, op, rgb = ps
    for i in range(reps + 20):
        t = [time.perf_counter()]
        uv, depth = G.project_point(xyz, intr, extr, W, H); t.append(time.perf_counter())
        vis = depth != 0; t.append(time.perf_counter())
        cov = G.compute_cov3d(scale, rot, vis); t.append(time.perf_counter())
        conic, radius, tiles = G.ewa_project(xyz, cov, intr, extr, uv, W, H, vis); t.append(time.perf_counter())
        ids, rng = G.sort_gaussian(uv, depth, W, H, radius, tiles); t.append(time.perf_counter())
        img = G.alpha_blending(uv, conic, op, rgb, ids, rng, 0.0, W, H); t.append(time.perf_counter())
        if i % 4 == 3: torch.cuda.synchronize()
        if i >= 20:
            for n, a, b in zip(names, t[:-1], t[1:]): acc[n].append(b - a)
    print("chain forward, host us per operator: " + "  ".join(f"{n} {statistics.median(v)*1e6:.1f}" for n, v in acc.items()))
per_op()
with torch.no_grad():
    print("(no_grad)", end=" ")
    per_op()
