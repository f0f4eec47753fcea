#!/usr/bin/env python
"""Time the blend forward / backward kernels of several library variants (tools/build_variants.py) on the same
input, each kernel alone, L2 flushed between launches, CUDA events on the launching stream; check every variant's
image and gradient pack against the first one.

    python tools/ab_blend.py [--workload cfg2] [--profile synthetic] [--reps 30] [name ...]
"""
import argparse
import ctypes
import glob
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gflow_b200 as G  # noqa: E402
from gflow_b200 import _build, capi  # noqa: E402
from gflow_b200.synthetic import CONFIGS, make_grad_image, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--profile", default="synthetic")
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--channels", type=int, default=3)
ap.add_argument("names", nargs="*")
a = ap.parse_args()

dev = torch.device("cuda:0")
N, W, H = CONFIGS[a.workload]
sc = make_scene(N, W, H, seed=0, profile=a.profile)
xyz, scale, rot, op, rgb, intr, extr = (t.to(dev) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, sc.intr, sc.extr))
with torch.no_grad():
    uv, depth = G.project_point(xyz, intr, extr, W, H)
    vis = depth != 0
    cov = G.compute_cov3d(scale, rot, vis)
    conic, radius, tiles = G.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
    ids, rng = G.sort_gaussian(uv, depth, W, H, radius, tiles)
K = ids.numel()
C = a.channels
feat = torch.cat([rgb, depth], dim=1)[:, :C].contiguous() if C <= 4 else None
Gimg = make_grad_image(C, W, H).to(dev)
opf = op.reshape(-1).contiguous()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

vdir = os.path.join(_build.LIB_DIR, "variants")
names = a.names or sorted(os.path.basename(p)[len("libgfb_"):-3] for p in glob.glob(os.path.join(vdir, "libgfb_*.so")))
if "base" in names:
    names.remove("base")
    names.insert(0, "base")
print(f"# {a.workload} ({N} Gaussians, {W}x{H}), profile {a.profile}, K = {K}, C = {C}, {a.reps} reps, L2 flushed")
ref = None
for name in names:
    lib = ctypes.CDLL(os.path.join(vdir, f"libgfb_{name}.so"))
    for sym in ("gfb_blend_pack_geometry", "gfb_blend_pack_feature", "gfb_alpha_blending_fwd", "gfb_alpha_blending_bwd"):
        fn = getattr(lib, sym)
        fn.restype, fn.argtypes = capi.SIGNATURES[sym]
    geom = torch.empty(max(K, 1) * 8, device=dev)
    fs = torch.empty(max(K, 1) * 4, device=dev)
    out = torch.empty(C, H, W, device=dev)
    fT = torch.empty(H, W, device=dev)
    nc = torch.empty(H, W, device=dev, dtype=torch.int32)
    gp = torch.zeros(N * 12, device=dev)
    assert lib.gfb_blend_pack_geometry(uv.data_ptr(), conic.data_ptr(), opf.data_ptr(), ids.data_ptr(), K, geom.data_ptr(), st) == 0
    assert lib.gfb_blend_pack_feature(feat.data_ptr(), C, 0, C, ids.data_ptr(), K, fs.data_ptr(), st) == 0

    def fwd():
        rc = lib.gfb_alpha_blending_fwd(geom.data_ptr(), fs.data_ptr(), K, rng.data_ptr(), C, 0, C, 0.0, W, H,
                                        out.data_ptr(), fT.data_ptr(), nc.data_ptr(), st)
        assert rc == 0, rc

    def bwd():
        rc = lib.gfb_alpha_blending_bwd(geom.data_ptr(), fs.data_ptr(), K, ids.data_ptr(), rng.data_ptr(), C, 0, C, 0.0,
                                        W, H, fT.data_ptr(), nc.data_ptr(), Gimg.data_ptr(), gp.data_ptr(), st)
        assert rc == 0, rc

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.mean(ts), statistics.median(ts), min(ts)

    fwd()
    tf = timeit(fwd)
    tb = timeit(bwd)
    gp.zero_()
    bwd()
    torch.cuda.synchronize()
    res = (out.clone(), gp.clone())
    if ref is None:
        ref = res
        d_img = d_gp = 0.0
    else:
        d_img = float((res[0] - ref[0]).abs().max() / ref[0].abs().max())
        d_gp = float((res[1] - ref[1]).abs().max() / ref[1].abs().max())
    print(f"{name:12s} fwd {tf[0]:7.2f} us (median {tf[1]:7.2f}, min {tf[2]:7.2f})   bwd {tb[0]:7.2f} us (median {tb[1]:7.2f}, "
          f"min {tb[2]:7.2f})   vs first: image {d_img:.2e} grad {d_gp:.2e}")
