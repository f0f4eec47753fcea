"""TEST-ONLY: dynamic operation counts of the kernels at a named workload, taken from the SIMT shim's counters
(warp collectives, lane-level atomics, CTA barriers, bytes staged by cp.async.bulk).  They are properties of the
algorithm on that input, not timings:

    python tests/simt/op_counts.py [cfg1|cfg2] [synthetic|gflow]            # per kernel group of the fused render step
    GFB_BWD_SPARSE=4 python tests/simt/op_counts.py cfg2 gflow               # the experimental backward variant
"""
import ctypes
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import emu  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
profile = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
N, W, H = CONFIGS[workload]
L = emu.load()
L.gfb_emu_stats.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_ulonglong * 8)()


def stats():
    L.gfb_emu_stats(ctypes.addressof(buf))
    return list(buf)


sc = make_scene(N, W, H, seed=0, profile=profile)
Gimg = make_grad_image(3, W, H)
p = emu.p
gx, gy = (W + 15) // 16, (H + 15) // 16
T = gx * gy
xyz, scale, rot, op, rgb = (t.contiguous() for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb))
intr, extr = sc.intr.contiguous(), sc.extr.contiguous()
uv, depth, conic, radius = emu.f32(N, 2), emu.f32(N, 1), emu.f32(N, 3), emu.i32(N, 1)
rect = torch.zeros(N * 8, dtype=torch.uint8)
ctrl = torch.zeros(L.gfb_render_control_bytes(W, H), dtype=torch.uint8)
rng = emu.i32(T, 2)
cap = 8 * N
keys, ids = torch.zeros(cap * 8, dtype=torch.uint8), emu.i32(cap)
geom, feat = torch.zeros(cap * 8), torch.zeros(cap * 4)
out, fT, nc = emu.f32(3, H, W), emu.f32(H, W), emu.i32(H, W)
K = ctypes.c_int64(-1)
L.gfb_emu_stats_reset()
t0 = time.time()
emu.ok(L.gfb_render_forward(p(xyz), p(scale), p(rot), p(op), p(rgb), 3, p(intr), p(extr), N, W, H, 0.0, 0.2, 1.3, p(uv), p(depth),
                            p(conic), p(radius), p(rect), p(ctrl), p(rng), cap, p(keys), p(ids), p(geom), p(feat), p(out), p(fT),
                            p(nc), ctypes.addressof(K), None), "forward")
fwd = stats()
gws = torch.zeros(L.gfb_render_grad_bytes(N) // 4)
d = [emu.f32(N, 3), emu.f32(N, 3), emu.f32(N, 4), emu.f32(N, 1), emu.f32(N, 3)]
L.gfb_emu_stats_reset()
emu.ok(L.gfb_render_backward(p(xyz), p(scale), p(rot), p(intr), p(extr), N, W, H, 3, 0.0, 0.2, 1.3, p(ids), p(rng), cap, p(geom),
                             p(feat), p(fT), p(nc), p(Gimg), p(gws), p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), None), "backward")
bwd = stats()
names = ["warp collectives", "lane-level atomics", "CTA barriers", "bulk-copy bytes", "CTAs", "threads"]
print(f"# {workload} ({N} Gaussians, {W}x{H}), profile {profile}, K = {K.value}, GFB_BWD_SPARSE={os.environ.get('GFB_BWD_SPARSE', '0')}"
      f"  (emulated in {time.time() - t0:.0f} s)")
print(f"{'':22s} {'forward (4 kernels)':>22s} {'backward (2 kernels)':>22s}")
for i, n in enumerate(names):
    print(f"{n:22s} {fwd[i]:22,d} {bwd[i]:22,d}")
