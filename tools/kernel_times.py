#!/usr/bin/env python
"""In-situ per-kernel durations of a render step (torch.profiler / CUPTI; no cache flush, no serialisation):
python tools/kernel_times.py [fused|chain] [steps] [workload] [profile] [flush]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import gflow_b200 as G  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "fused"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N, W, H = CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "cfg2"]
prof_name = sys.argv[4] if len(sys.argv) > 4 else "synthetic"
flush = len(sys.argv) > 5 and sys.argv[5] == "flush"
dev = torch.device("cuda:0")
sc = make_scene(N, W, H, seed=0, profile=prof_name)
ps = [t.to(dev).requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb)]
intr, extr = sc.intr.to(dev), sc.extr.to(dev).requires_grad_(True)
Gimg = make_grad_image(3, W, H).to(dev)
fn = G.rasterization if path == "fused" else G.rasterization_unfused
buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    for p in ps:
        p.grad = None
    extr.grad = None
    img = fn(*ps, intr, extr, W, H, 0.0)
    img.backward(Gimg)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        if flush:
            buf.fill_(1)
        step()
    torch.cuda.synchronize()
tot = collections.OrderedDict()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        d = tot.setdefault(name, [0.0, 0])
        d[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        d[1] += 1
print(f"# in-situ kernel times, path={path} workload={N}x{W}x{H} profile={prof_name} flush={flush}, avg per step over {steps} steps")
s = 0.0
for name, (t, n) in tot.items():
    if "FillFunctor<unsigned char" in name:
        continue
    print(f"{t / steps:9.2f} us/step  x{n / steps:4.1f}  {name}")
    s += t / steps
print(f"{s:9.2f} us/step  total GPU busy")
# timeline of the last step: start offset, duration and the idle gap before each GPU activity
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
             key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(evs) if "preprocess_kernel" in e.name or "project_point_fwd" in e.name]
if len(starts) >= 2:
    a, b = starts[-2], starts[-1]
    t0 = evs[a].time_range.start
    prev_end = t0
    print("# timeline of one step (us from the first kernel): start  dur  gap_before  name")
    for e in evs[a:b]:
        st, en = e.time_range.start, e.time_range.end
        nm = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:50]
        print(f"  {st - t0:8.1f} {en - st:7.1f} {st - prev_end:7.1f}  {nm}")
        prev_end = max(prev_end, en)

