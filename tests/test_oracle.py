"""CPU tests of the oracle itself (no GPU, no CUDA calls).

The oracle is the checker for the CUDA path, so it is pinned three ways:
  * two independent restatements (plain C, PyTorch) must agree -- bit for bit on the
    geometry / integer outputs, and the C analytic backward against PyTorch autograd;
  * the PyTorch restatement passes torch.autograd.gradcheck in float64;
  * both reproduce the committed golden vectors (tests/golden/*.npz).
PARITY UNPINNED w.r.t. real MSplat: the reference ships no vectors.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import assert_close
from gflow_b200.synthetic import make_grad_image, make_scene
from oracle import c_oracle as C
from oracle import splat_ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


@pytest.mark.parametrize("seed,profile", [(0, "synthetic"), (1, "gflow")])
def test_c_and_torch_oracle_agree_forward(seed, profile):
    sc = make_scene(2000, 200, 136, seed=seed, profile=profile)
    W, H = sc.W, sc.H
    uv_r, d_r = R.project_point(sc.xyz, sc.intr, sc.extr, W, H)
    uv_c, d_c = C.project_point(sc.xyz, sc.intr, sc.extr, W, H)
    assert torch.equal(uv_r, uv_c) and torch.equal(d_r, d_c)
    vis = d_r != 0
    assert 0 < int(vis.sum()) < 2000  # culling is exercised
    cov_r, cov_c = R.compute_cov3d(sc.scale, sc.rotate, vis), C.compute_cov3d(sc.scale, sc.rotate, vis)
    assert torch.equal(cov_r, cov_c)
    con_r, rad_r, t_r = R.ewa_project(sc.xyz, cov_r, sc.intr, sc.extr, uv_r, W, H, vis)
    con_c, rad_c, t_c = C.ewa_project(sc.xyz, cov_r, sc.intr, sc.extr, uv_r, W, H, vis)
    assert torch.equal(rad_r, rad_c) and torch.equal(t_r, t_c) and torch.equal(con_r, con_c)
    ids_r, rng_r = R.sort_gaussian(uv_r, d_r, W, H, rad_r, t_r)
    ids_c, rng_c = C.sort_gaussian(uv_r, d_r, W, H, rad_r, t_r)
    assert torch.equal(ids_r, ids_c) and torch.equal(rng_r, rng_c)
    assert ids_r.numel() == int(t_r.sum())
    for feat, bg in ((sc.rgb, 0.0), (d_r, 0.5)):
        img_r, fT_r, nc_r = R.alpha_blending(uv_r, con_r, sc.opacity, feat, ids_r, rng_r, bg, W, H, return_aux=True)
        img_c, fT_c, nc_c = C.alpha_blending(uv_r, con_r, sc.opacity, feat, ids_r, rng_r, bg, W, H, return_aux=True)
        assert_close(img_c, img_r, 1e-5, "image")
        assert_close(fT_c, fT_r, 1e-5, "final_T")
        assert int((nc_r != nc_c).sum()) <= 2


def test_c_backward_matches_autograd_float64_chain():
    sc = make_scene(1500, 160, 120, seed=5)
    W, H = sc.W, sc.H
    G = make_grad_image(3, W, H)
    xs = [t.double().requires_grad_(True) for t in (sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, sc.intr, sc.extr)]
    xyz, scale, rot, op, rgb, intr, extr = xs
    uv, d = R.project_point(xyz, intr, extr, W, H)
    v = d != 0
    cov = R.compute_cov3d(scale, rot, v)
    con, rad, tl = R.ewa_project(xyz, cov, intr, extr, uv, W, H, v)
    ids, rng = R.sort_gaussian(uv, d, W, H, rad, tl)
    img = R.alpha_blending(uv, con, op, rgb, ids, rng, 0.3, W, H)
    (img * G.double()).sum().backward()
    img_c, gc, info = C.render_step_fwd_bwd(sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, sc.intr, sc.extr, 0.3,
                                            W, H, G)
    assert info["K"] == ids.numel()
    assert_close(img_c, img.float(), 1e-4, "image")
    for name, g in zip(["xyz", "scale", "rotate", "opacity", "feature", "intr", "extr"], xs):
        assert_close(gc[name], g.grad.float(), 1e-3, "grad " + name)


def test_ewa_frustum_clamp_backward():
    """Points outside 1.3 x the frustum (visible=None) take the clamped-Jacobian branch."""
    gen = torch.Generator().manual_seed(3)
    N, W, H = 400, 128, 96
    xyz = torch.cat([(torch.rand(N, 2, generator=gen) - 0.5) * 8.0, 1.0 + torch.rand(N, 1, generator=gen)], dim=1)
    sc = make_scene(N, W, H, seed=4)
    intr, extr = sc.intr, sc.extr
    cov = C.compute_cov3d(sc.scale * 20, sc.rotate, None)
    uv = torch.rand(N, 2, generator=gen) * torch.tensor([W, H])
    con_c, rad_c, t_c = C.ewa_project(xyz, cov, intr, extr, uv, W, H, None)
    x64 = [t.double().requires_grad_(True) for t in (xyz, cov, intr, extr)]
    con_r, rad_r, t_r = R.ewa_project(x64[0], x64[1], x64[2], x64[3], uv.double(), W, H, None)
    tx = (xyz @ extr[:, :3].T + extr[:, 3])
    clamped = (tx[:, 0] / tx[:, 2]).abs() > 1.3 * W / (2 * intr[0])
    assert int(clamped.sum()) > 20, "test scene must hit the clamp"
    g = torch.randn(N, 3, generator=gen)
    (con_r * g.double()).sum().backward()
    d_xyz, d_cov, d_intr, d_extr = C.ewa_project_bwd(xyz, cov, intr, extr, uv, W, H, None, g)
    assert_close(con_c, con_r.float(), 1e-5, "conic")
    assert torch.equal(rad_c, rad_r)
    assert_close(d_xyz, x64[0].grad.float(), 1e-3, "d_xyz")
    assert_close(d_cov, x64[1].grad.float(), 1e-3, "d_cov3d")
    assert_close(d_intr[:2], x64[2].grad.float()[:2], 1e-3, "d_intr")
    assert_close(d_extr, x64[3].grad.float(), 1e-3, "d_extr")


def test_gradcheck_geometry_float64():
    gen = torch.Generator().manual_seed(0)
    N, W, H = 6, 64, 48
    sc = make_scene(N, W, H, seed=9, outside_frac=0.0)
    xyz, intr, extr = (t.double().requires_grad_(True) for t in (sc.xyz, sc.intr, sc.extr))
    assert torch.autograd.gradcheck(lambda a, b, c: R.project_point(a, b, c, W, H), (xyz, intr, extr), eps=1e-6,
                                    atol=1e-6, rtol=1e-4)
    scale = (sc.scale.double() * 50).requires_grad_(True)
    rot = sc.rotate.double().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda s, q: R.compute_cov3d(s, q, None), (scale, rot), eps=1e-6, atol=1e-6,
                                    rtol=1e-4)
    cov = R.compute_cov3d(scale, rot, None).detach().requires_grad_(True)
    uv = R.project_point(xyz, intr, extr, W, H)[0].detach()
    assert torch.autograd.gradcheck(lambda p, s, i, e: R.ewa_project(p, s, i, e, uv, W, H, None)[0],
                                    (xyz, cov, intr, extr), eps=1e-6, atol=1e-5, rtol=1e-3)
    shs = torch.randn(5, 3, 16, generator=gen, dtype=torch.float64).requires_grad_(True)
    dirs = torch.randn(5, 3, generator=gen, dtype=torch.float64).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda s, d: R.compute_sh(s, d), (shs, dirs), eps=1e-6, atol=1e-6, rtol=1e-4)


def test_gradcheck_alpha_blending_float64():
    # opacities < 0.99 and well-separated alphas: the (transparent) 0.99 clamp and the 1/255 /
    # 1e-4 thresholds are not crossed by the finite-difference step.
    gen = torch.Generator().manual_seed(1)
    N, W, H = 12, 32, 16
    uv = (torch.rand(N, 2, generator=gen, dtype=torch.float64) * torch.tensor([W, H])).requires_grad_(True)
    s = 2.0 + 3.0 * torch.rand(N, generator=gen, dtype=torch.float64)
    conic = torch.stack([1 / s ** 2, 0.02 * torch.randn(N, generator=gen, dtype=torch.float64), 1 / s ** 2], 1)
    conic = conic.requires_grad_(True)
    opacity = (0.2 + 0.5 * torch.rand(N, 1, generator=gen, dtype=torch.float64)).requires_grad_(True)
    feat = torch.rand(N, 3, generator=gen, dtype=torch.float64).requires_grad_(True)
    depth = 1.0 + torch.rand(N, 1, generator=gen)
    radius = torch.full((N, 1), 20, dtype=torch.int32)
    tiles = torch.full((N, 1), 2, dtype=torch.int32)
    ids, rng = R.sort_gaussian(uv.detach().float(), depth, W, H, radius, tiles)

    def f(u, c, o, ft):
        return R.alpha_blending(u, c, o, ft, ids, rng, 0.4, W, H)

    assert torch.autograd.gradcheck(f, (uv, conic, opacity, feat), eps=1e-6, atol=1e-5, rtol=1e-3)


def test_sh_c_matches_torch():
    gen = torch.Generator().manual_seed(2)
    for K in (1, 4, 9, 16):
        shs = torch.randn(50, 3, K, generator=gen)
        dirs = torch.randn(50, 3, generator=gen)
        vis = torch.rand(50, 1, generator=gen) > 0.2
        s64, d64 = shs.double().requires_grad_(True), dirs.double().requires_grad_(True)
        out = R.compute_sh(s64, d64, vis)
        g = torch.randn(50, 3, generator=gen)
        (out * g.double()).sum().backward()
        assert_close(C.compute_sh(shs, dirs, vis), out.float(), 1e-5, f"sh K={K}")
        d_shs, d_dirs = C.compute_sh_bwd(shs, dirs, vis, g)
        assert_close(d_shs, s64.grad.float(), 1e-4, "d_shs")
        if K > 1:
            assert_close(d_dirs, d64.grad.float(), 1e-4, "d_dirs")


def test_sort_edge_cases():
    W, H = 40, 40  # 3x3 tiles
    T = 9
    e = torch.zeros(0, 2)
    ids, rng = C.sort_gaussian(e, torch.zeros(0, 1), W, H, torch.zeros(0, 1, dtype=torch.int32),
                               torch.zeros(0, 1, dtype=torch.int32))
    assert ids.numel() == 0 and rng.shape == (T, 2) and int(rng.abs().sum()) == 0
    # equal depths keep Gaussian-id order; a Gaussian covering every tile appears in all of them
    uv = torch.tensor([[20.0, 20.0], [5.0, 5.0], [5.0, 5.0], [5.0, 5.0]])
    depth = torch.tensor([[2.0], [1.0], [1.0], [0.5]])
    radius = torch.tensor([[100], [3], [3], [3]], dtype=torch.int32)
    tiles = torch.tensor([[9], [1], [1], [1]], dtype=torch.int32)
    for mod in (C, R):
        ids, rng = mod.sort_gaussian(uv, depth, W, H, radius, tiles)
        assert ids.tolist()[:4] == [3, 1, 2, 0]
        assert rng[0].tolist() == [0, 4] and ids.numel() == 12
        assert all(rng[t].tolist() == [3 + t, 4 + t] for t in range(1, 9))


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "splat_*.npz"))))
def test_oracles_reproduce_golden_splat(name):
    g = _load(name)
    W, H, bg = g["W"], g["H"], g["bg"]
    for mod in (C, R):
        uv, depth = mod.project_point(g["xyz"], g["intr"], g["extr"], W, H)
        assert torch.equal(uv, g["uv"]) and torch.equal(depth, g["depth"])
        vis = depth != 0
        cov = mod.compute_cov3d(g["scale"], g["rotate"], vis)
        assert torch.equal(cov, g["cov3d"])
        conic, radius, tiles = mod.ewa_project(g["xyz"], cov, g["intr"], g["extr"], uv, W, H, vis)
        assert torch.equal(conic, g["conic"]) and torch.equal(radius, g["radius"]) and torch.equal(tiles, g["tiles"])
        ids, rng = mod.sort_gaussian(uv, depth, W, H, radius, tiles)
        assert torch.equal(ids, g["ids"]) and torch.equal(rng, g["tile_range"])
        img = mod.alpha_blending(uv, conic, g["opacity"], g["feature"], ids, rng, bg, W, H)
        assert_close(img, g["img"], 1e-5, "image")
    d_uv, d_conic, d_op, d_f = C.alpha_blending_bwd(g["uv"], g["conic"], g["opacity"], g["feature"], g["ids"],
                                                    g["tile_range"], bg, W, H, g["final_T"], g["n_contrib"],
                                                    g["g_img"])
    for a, k in ((d_uv, "d_uv"), (d_conic, "d_conic"), (d_op, "d_opacity"), (d_f, "d_feature")):
        assert_close(a, g[k], 1e-5, k)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_oracles_reproduce_golden_sh(deg):
    g = _load(f"sh_deg{deg}.npz")
    for mod in (C, R):
        assert_close(mod.compute_sh(g["shs"], g["dirs"], g["visible"]), g["out"], 1e-5, "sh")
    d_shs, d_dirs = C.compute_sh_bwd(g["shs"], g["dirs"], g["visible"], g["g_out"])
    assert_close(d_shs, g["d_shs"], 1e-6, "d_shs")
    assert_close(d_dirs, g["d_dirs"], 1e-5, "d_dirs")
