"""Generate the committed golden vectors from the oracle (run from the repo root):

    python tests/golden/make_golden.py

PARITY UNPINNED: /root/reference holds no golden vectors for this path and msplat itself
is not installable here, so these fixtures freeze the oracle's behaviour
(oracle/splat_oracle.c, cross-checked against oracle/splat_ref.py) -- they pin the CUDA
path and the oracle to each other across rounds, not to real MSplat.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gflow_b200.synthetic import make_grad_image, make_scene  # noqa: E402
from oracle import c_oracle as C  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make(name, N, W, H, seed, profile, bg, Cfeat):
    sc = make_scene(N, W, H, seed=seed, profile=profile, bg=bg)
    uv, depth = C.project_point(sc.xyz, sc.intr, sc.extr, W, H)
    vis = depth != 0
    cov3d = C.compute_cov3d(sc.scale, sc.rotate, vis)
    conic, radius, tiles = C.ewa_project(sc.xyz, cov3d, sc.intr, sc.extr, uv, W, H, vis)
    ids, rng = C.sort_gaussian(uv, depth, W, H, radius, tiles)
    gen = torch.Generator().manual_seed(seed + 100)
    feat = torch.rand(N, Cfeat, generator=gen)
    img, fT, nc = C.alpha_blending(uv, conic, sc.opacity, feat, ids, rng, bg, W, H, return_aux=True)
    G = make_grad_image(Cfeat, W, H, seed=seed + 1)
    d_uv, d_conic, d_op, d_f = C.alpha_blending_bwd(uv, conic, sc.opacity, feat, ids, rng, bg, W, H, fT, nc, G)
    d_xyz_e, d_cov, d_intr_e, d_extr_e = C.ewa_project_bwd(sc.xyz, cov3d, sc.intr, sc.extr, uv, W, H, vis, d_conic)
    d_s, d_q = C.compute_cov3d_bwd(sc.scale, sc.rotate, vis, d_cov)
    g_depth = torch.randn(N, 1, generator=gen)
    d_xyz_p, d_intr_p, d_extr_p = C.project_point_bwd(sc.xyz, sc.intr, sc.extr, W, H, d_uv, g_depth)
    arrs = dict(
        W=W, H=H, bg=bg, xyz=sc.xyz, scale=sc.scale, rotate=sc.rotate, opacity=sc.opacity, feature=feat,
        intr=sc.intr, extr=sc.extr, uv=uv, depth=depth, cov3d=cov3d, conic=conic, radius=radius, tiles=tiles,
        ids=ids, tile_range=rng, img=img, final_T=fT, n_contrib=nc, g_img=G, g_depth=g_depth, d_uv=d_uv,
        d_conic=d_conic, d_opacity=d_op, d_feature=d_f, d_xyz_ewa=d_xyz_e, d_cov3d=d_cov, d_intr_ewa=d_intr_e,
        d_extr_ewa=d_extr_e, d_scale=d_s, d_rotate=d_q, d_xyz_proj=d_xyz_p, d_intr_proj=d_intr_p,
        d_extr_proj=d_extr_p,
    )
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(name, "N", N, "K", ids.numel(), "visible", int(vis.sum()))


def make_sh(name, N, Cf, K, seed):
    gen = torch.Generator().manual_seed(seed)
    shs = torch.randn(N, Cf, K, generator=gen)
    dirs = torch.randn(N, 3, generator=gen) * 3.0
    vis = torch.rand(N, 1, generator=gen) > 0.1
    g = torch.randn(N, Cf, generator=gen)
    out = C.compute_sh(shs, dirs, vis)
    d_shs, d_dirs = C.compute_sh_bwd(shs, dirs, vis, g)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), shs=shs.numpy(), dirs=dirs.numpy(), visible=vis.numpy(),
                        g_out=g.numpy(), out=out.numpy(), d_shs=d_shs.numpy(), d_dirs=d_dirs.numpy())
    print(name)


if __name__ == "__main__":
    make("splat_small_c3", 300, 70, 52, seed=11, profile="synthetic", bg=0.0, Cfeat=3)
    make("splat_small_c1_bg", 300, 64, 48, seed=12, profile="gflow", bg=0.33, Cfeat=1)
    make("splat_small_c5", 200, 48, 40, seed=13, profile="synthetic", bg=1.0, Cfeat=5)
    for deg, K in enumerate((1, 4, 9, 16)):
        make_sh(f"sh_deg{deg}", 64, 3, K, seed=20 + deg)
