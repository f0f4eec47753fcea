#!/usr/bin/env python
"""Eager-PyTorch-on-CUDA proxy baseline: the oracle's differentiable PyTorch restatement
of the render step (oracle/splat_ref.py) executed with CUDA tensors -- what a pure-PyTorch port of the path costs on the
same GPU.  It is NOT MSplat (which is not installable here) and is labelled as a proxy wherever it is printed.

    python tools/eager_proxy.py [cfg2] [synthetic] [steps]
"""
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import splat_ref as R  # noqa: E402

spec = importlib.util.spec_from_file_location("gfb_synthetic_standalone", os.path.join(ROOT, "gflow_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
sys.modules["gfb_synthetic_standalone"] = syn
spec.loader.exec_module(syn)

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
profile = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N, W, H = syn.CONFIGS[workload]
dev = torch.device("cuda:0")
sc = syn.make_scene(N, W, H, seed=0, profile=profile)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = syn.make_grad_image(3, W, H).to(dev)


def step():
    for p in ps:
        p.grad = None
    img, _ = R.render_step(*ps, intr, extr, sc.bg, W, H)
    img.backward(Gimg)


step()  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"value": steps / dt, "unit": "iters/s", "kind": "proxy", "ms_per_step": 1e3 * dt / steps, "steps": steps,
                  "what": "oracle/splat_ref.py (eager PyTorch, autograd) with CUDA tensors on the same GPU: a labelled PROXY for "
                          "a GPU comparator, not MSplat"}))
