#!/usr/bin/env python
"""Per-CTA timeline of the blend backward (variant library built with -DGFB_BLEND_TRACE, tools/build_variants.py):
which SM ran each tile, when it started and ended.  Prints how evenly the tiles were spread over the SMs and how
long the SMs sat idle before the last CTA finished."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gflow_b200 as G  # noqa: E402
from gflow_b200 import _build, capi  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
profile = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
variant = sys.argv[3] if len(sys.argv) > 3 else "trace"
dev = torch.device("cuda:0")
N, W, H = CONFIGS[workload]
sc = make_scene(N, W, H, seed=0, profile=profile)
xyz, scale, rot, op, rgb, intr, extr = (t.to(dev) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, sc.intr, sc.extr))
with torch.no_grad():
    uv, depth = G.project_point(xyz, intr, extr, W, H)
    vis = depth != 0
    cov = G.compute_cov3d(scale, rot, vis)
    conic, radius, tiles = G.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
    ids, rng = G.sort_gaussian(uv, depth, W, H, radius, tiles)
K = ids.numel()
C = 3
Gimg = make_grad_image(C, W, H).to(dev)
opf = op.reshape(-1).contiguous()
st = torch.cuda.current_stream().cuda_stream
lib = ctypes.CDLL(os.path.join(_build.LIB_DIR, "variants", f"libgfb_{variant}.so"))
for sym in ("gfb_blend_pack_geometry", "gfb_blend_pack_feature", "gfb_alpha_blending_fwd", "gfb_alpha_blending_bwd"):
    fn = getattr(lib, sym)
    fn.restype, fn.argtypes = capi.SIGNATURES[sym]
lib.gfb_debug_blend_trace.restype = ctypes.c_int
lib.gfb_debug_blend_trace.argtypes = [ctypes.c_void_p]
geom = torch.empty(max(K, 1) * 8, device=dev)
fs = torch.empty(max(K, 1) * 4, device=dev)
out = torch.empty(C, H, W, device=dev)
fT = torch.empty(H, W, device=dev)
nc = torch.empty(H, W, device=dev, dtype=torch.int32)
gp = torch.zeros(N * 12, device=dev)
lib.gfb_blend_pack_geometry(uv.data_ptr(), conic.data_ptr(), opf.data_ptr(), ids.data_ptr(), K, geom.data_ptr(), st)
lib.gfb_blend_pack_feature(rgb.data_ptr(), C, 0, C, ids.data_ptr(), K, fs.data_ptr(), st)
lib.gfb_alpha_blending_fwd(geom.data_ptr(), fs.data_ptr(), K, rng.data_ptr(), C, 0, C, 0.0, W, H, out.data_ptr(), fT.data_ptr(),
                           nc.data_ptr(), st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    flush.fill_(1)
    lib.gfb_alpha_blending_bwd(geom.data_ptr(), fs.data_ptr(), K, ids.data_ptr(), rng.data_ptr(), C, 0, C, 0.0, W, H,
                               fT.data_ptr(), nc.data_ptr(), Gimg.data_ptr(), gp.data_ptr(), st)
torch.cuda.synchronize()
buf = np.zeros(3 * 8192, dtype=np.uint64)
assert lib.gfb_debug_blend_trace(buf.ctypes.data) == 0
T = rng.shape[0]
n_t = (rng[:, 1] - rng[:, 0]).cpu().numpy()
tr = buf.reshape(-1, 3)[:T].astype(np.int64)
live = n_t > 0
smid, t0, t1 = tr[live, 0], tr[live, 1], tr[live, 2]
base = t0.min()
t0 = (t0 - base) / 1e3
t1 = (t1 - base) / 1e3
print(f"# {workload} {profile}: {live.sum()} non-empty tiles of {T}, K = {K}; kernel span {t1.max():.1f} us (first CTA start to last CTA end)")
nsm = int(smid.max()) + 1
cnt = np.bincount(smid, minlength=nsm)
print(f"SMs used {np.count_nonzero(cnt)}; CTAs per SM: min {cnt[cnt > 0].min()} mean {cnt[cnt > 0].mean():.2f} max {cnt.max()}; histogram {np.bincount(cnt)[1:].tolist()} (index = count-1 from 1)")
end_sm = np.array([t1[smid == s].max() if cnt[s] else 0 for s in range(nsm)])
work_sm = np.array([n_t[live][smid == s].sum() if cnt[s] else 0 for s in range(nsm)])
u = cnt > 0
print(f"per-SM last CTA end: min {end_sm[u].min():.1f} mean {end_sm[u].mean():.1f} max {end_sm[u].max():.1f} us")
print(f"per-SM records: min {work_sm[u].min()} mean {work_sm[u].mean():.0f} max {work_sm[u].max()}  corr(end, records) = {np.corrcoef(end_sm[u], work_sm[u])[0, 1]:.3f}  corr(end, ctas) = {np.corrcoef(end_sm[u], cnt[u])[0, 1]:.3f}")
dur = t1 - t0
print(f"CTA start: p50 {np.percentile(t0, 50):.1f} p90 {np.percentile(t0, 90):.1f} p99 {np.percentile(t0, 99):.1f} max {t0.max():.1f} us")
print(f"CTA duration: p10 {np.percentile(dur, 10):.1f} p50 {np.percentile(dur, 50):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f} us; late starters (start > 5 us): {(t0 > 5).sum()} with mean duration {dur[t0 > 5].mean() if (t0 > 5).any() else 0:.1f} us")
# concurrency over time
ts = np.linspace(0, t1.max(), 21)
conc = [int(((t0 <= x) & (t1 > x)).sum()) for x in ts]
print("resident CTAs over time (20 steps):", conc)
