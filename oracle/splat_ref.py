"""CPU restatement (PyTorch, differentiable) of the `msplat` operator surface GFlow calls.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.
The product path (``gflow_b200``) never does, and has no CPU fallback.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
``msplat`` (github.com/pointrix-project/msplat, version unpinned by the reference:
``/root/reference/README.md:27``); it is absent from ``/root/reference``, not
installable here, and the reference ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c).  The functions below therefore follow
  (1) the call-site contracts of ``/root/reference/gflow/utils/render.py:21-154`` and
      ``/root/reference/gflow/trainer.py:29-42,115-121,953-955`` (shapes, dtypes,
      conventions), and
  (2) the published algorithm of the 3DGS rasteriser lineage MSplat derives from
      (SURVEY.md Appendix A, items flagged [R] there are recalled, not verified).

Every function is written with plain tensor ops in a fixed operation order (one
rounding per multiply / add, no fused multiply-add) so that the float32 forward
results of the per-Gaussian geometry are reproducible bit for bit by the C
restatement (``oracle/splat_oracle.c``) and by the CUDA kernels.

Works in float32 and float64 (float64 is used for ``torch.autograd.gradcheck``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

TILE = 16
ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.99
T_EPS = 1e-4
COV_BLUR = 0.3
FRUSTUM_CLAMP = 1.3


# --------------------------------------------------------------------------- helpers
def _cam_point(xyz, extr):
    """p_c = R p + t with the fixed order ((e0*x + e1*y) + e2*z) + e3.

    extr is the world->camera [R|t] 3x4 matrix GFlow builds in
    /root/reference/gflow/trainer.py:115-121.
    """
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    e = extr.reshape(-1)
    xc = ((e[0] * x + e[1] * y) + e[2] * z) + e[3]
    yc = ((e[4] * x + e[5] * y) + e[6] * z) + e[7]
    zc = ((e[8] * x + e[9] * y) + e[10] * z) + e[11]
    return xc, yc, zc


# --------------------------------------------------------------------------- a1
def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):
    """World -> pixel projection with near / frustum culling.

    Call sites: /root/reference/gflow/utils/render.py:21-24,116-119 and
    /root/reference/gflow/trainer.py:955.  intr = (fx, fy, cx, cy)
    (trainer.py:40); culled points return uv = 0, depth = 0 exactly
    (render.py:29 uses ``depth != 0`` as the visibility mask).
    Returns uv (N,2), depth (N,1).
    """
    fx, fy, cx, cy = intr[0], intr[1], intr[2], intr[3]
    xc, yc, zc = _cam_point(xyz, extr)
    near_ok = zc > nearest
    zs = torch.where(near_ok, zc, torch.ones_like(zc))
    u = (fx * xc) / zs + cx
    v = (fy * yc) / zs + cy
    xn = u / (0.5 * W) - 1.0
    yn = v / (0.5 * H) - 1.0
    ok = near_ok & (xn.abs() <= extent) & (yn.abs() <= extent)
    zero = torch.zeros_like(u)
    uv = torch.stack([torch.where(ok, u, zero), torch.where(ok, v, zero)], dim=1)
    depth = torch.where(ok, zc, zero).unsqueeze(1)
    return uv, depth


# --------------------------------------------------------------------------- a3
def compute_cov3d(scale, rotate, visible=None):
    """Sigma = R(q) diag(s^2) R(q)^T, upper triangle (xx, xy, xz, yy, yz, zz).

    Call site: /root/reference/gflow/utils/render.py:37-41.  Quaternion is
    (w, x, y, z) and already normalised by the caller (trainer.py:66,723).
    Invisible Gaussians give zeros (and zero gradients).
    """
    w, x, y, z = rotate[:, 0], rotate[:, 1], rotate[:, 2], rotate[:, 3]
    sx, sy, sz = scale[:, 0], scale[:, 1], scale[:, 2]
    r00 = 1.0 - 2.0 * (y * y + z * z)
    r01 = 2.0 * (x * y - w * z)
    r02 = 2.0 * (x * z + w * y)
    r10 = 2.0 * (x * y + w * z)
    r11 = 1.0 - 2.0 * (x * x + z * z)
    r12 = 2.0 * (y * z - w * x)
    r20 = 2.0 * (x * z - w * y)
    r21 = 2.0 * (y * z + w * x)
    r22 = 1.0 - 2.0 * (x * x + y * y)
    m00, m01, m02 = r00 * sx, r01 * sy, r02 * sz
    m10, m11, m12 = r10 * sx, r11 * sy, r12 * sz
    m20, m21, m22 = r20 * sx, r21 * sy, r22 * sz
    c00 = (m00 * m00 + m01 * m01) + m02 * m02
    c01 = (m00 * m10 + m01 * m11) + m02 * m12
    c02 = (m00 * m20 + m01 * m21) + m02 * m22
    c11 = (m10 * m10 + m11 * m11) + m12 * m12
    c12 = (m10 * m20 + m11 * m21) + m12 * m22
    c22 = (m20 * m20 + m21 * m21) + m22 * m22
    cov = torch.stack([c00, c01, c02, c11, c12, c22], dim=1)
    if visible is not None:
        cov = torch.where(visible.reshape(-1, 1), cov, torch.zeros_like(cov))
    return cov


# --------------------------------------------------------------------------- a4
def _tile_rect(u, v, radius_f, W, H):
    """Tile rectangle on the 16x16 grid, 3DGS ``getRect`` rule."""
    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    x0 = torch.clamp(torch.trunc((u - radius_f) / TILE), 0, gx)
    y0 = torch.clamp(torch.trunc((v - radius_f) / TILE), 0, gy)
    x1 = torch.clamp(torch.trunc(((u + radius_f) + (TILE - 1)) / TILE), 0, gx)
    y1 = torch.clamp(torch.trunc(((v + radius_f) + (TILE - 1)) / TILE), 0, gy)
    return x0.to(torch.int64), y0.to(torch.int64), x1.to(torch.int64), y1.to(torch.int64)


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):
    """EWA splat: Sigma' = J W Sigma W^T J^T + 0.3 I, conic = Sigma'^-1.

    Call site: /root/reference/gflow/utils/render.py:44-49.  Returns
    conic (N,3) as (a, b, c) of [[a,b],[b,c]] (render.py:95-96 uses (1,0,1) as the
    identity), radius (N,1) int32 = ceil(3 sqrt(lambda_max)), tiles_touched (N,1)
    int32.  A Gaussian that is invisible, has det == 0 or touches no tile yields
    zeros in all three outputs.
    """
    fx, fy = intr[0], intr[1]
    e = extr.reshape(-1)
    tx, ty, tz = _cam_point(xyz, extr)
    vis = torch.ones_like(tz, dtype=torch.bool) if visible is None else visible.reshape(-1)
    tzs = torch.where(vis, tz, torch.ones_like(tz))
    # the clamp limits are constants w.r.t. autograd (fx enters only through J)
    limx = FRUSTUM_CLAMP * (float(W) / (2.0 * fx.detach()))
    limy = FRUSTUM_CLAMP * (float(H) / (2.0 * fy.detach()))
    txc = torch.minimum(limx, torch.maximum(-limx, tx / tzs)) * tzs
    tyc = torch.minimum(limy, torch.maximum(-limy, ty / tzs)) * tzs
    j00 = fx / tzs
    j02 = -(fx * txc) / (tzs * tzs)
    j11 = fy / tzs
    j12 = -(fy * tyc) / (tzs * tzs)
    # T = J * R  (2x3)
    t00 = j00 * e[0] + j02 * e[8]
    t01 = j00 * e[1] + j02 * e[9]
    t02 = j00 * e[2] + j02 * e[10]
    t10 = j11 * e[4] + j12 * e[8]
    t11 = j11 * e[5] + j12 * e[9]
    t12 = j11 * e[6] + j12 * e[10]
    s00, s01, s02, s11, s12, s22 = (cov3d[:, i] for i in range(6))
    # rows of Sigma * T_i
    a0 = (s00 * t00 + s01 * t01) + s02 * t02
    a1 = (s01 * t00 + s11 * t01) + s12 * t02
    a2 = (s02 * t00 + s12 * t01) + s22 * t02
    b0 = (s00 * t10 + s01 * t11) + s02 * t12
    b1 = (s01 * t10 + s11 * t11) + s12 * t12
    b2 = (s02 * t10 + s12 * t11) + s22 * t12
    ca = ((t00 * a0 + t01 * a1) + t02 * a2) + COV_BLUR
    cb = (t10 * a0 + t11 * a1) + t12 * a2
    cc = ((t10 * b0 + t11 * b1) + t12 * b2) + COV_BLUR
    det = ca * cc - cb * cb
    ok = vis & (det != 0)
    dinv = 1.0 / torch.where(ok, det, torch.ones_like(det))
    mid = 0.5 * (ca + cc)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius_f = torch.ceil(3.0 * torch.sqrt(lam.detach()))
    radius_f = torch.where(ok, radius_f, torch.zeros_like(radius_f))
    x0, y0, x1, y1 = _tile_rect(uv[:, 0].detach(), uv[:, 1].detach(), radius_f, W, H)
    tiles = (x1 - x0) * (y1 - y0)
    tiles = torch.where(ok, tiles, torch.zeros_like(tiles))
    ok = ok & (tiles > 0)
    zero = torch.zeros_like(ca)
    conic = torch.stack(
        [torch.where(ok, cc * dinv, zero), torch.where(ok, -cb * dinv, zero), torch.where(ok, ca * dinv, zero)],
        dim=1,
    )
    radius = torch.where(ok, radius_f, torch.zeros_like(radius_f)).to(torch.int32).unsqueeze(1)
    tiles = torch.where(ok, tiles, torch.zeros_like(tiles)).to(torch.int32).unsqueeze(1)
    return conic, radius, tiles


# --------------------------------------------------------------------------- a5
def sort_gaussian(uv, depth, W, H, radius, tiles_touched):
    """Duplicate each Gaussian per touched tile, order by (tile, depth bits), stable.

    Call site: /root/reference/gflow/utils/render.py:52-54.  Returns
    gaussian_ids_sorted (K,) int32 and tile_range (T,2) int32 ([start,end), zeros
    for empty tiles); tiles are row-major, tile = ty * ceil(W/16) + tx.
    Equal (tile, depth) keys keep Gaussian-id order (stable sort on a
    Gaussian-major emission).
    """
    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    T = gx * gy
    u = uv[:, 0].detach().to(torch.float32)
    v = uv[:, 1].detach().to(torch.float32)
    r = radius.reshape(-1)
    tt = tiles_touched.reshape(-1)
    x0, y0, x1, y1 = (a.cpu().numpy() for a in _tile_rect(u, v, r.to(torch.float32), W, H))
    live = ((r > 0) & (tt > 0)).cpu().numpy()
    dbits = depth.detach().reshape(-1).to(torch.float32).contiguous().cpu().numpy().view(np.uint32)
    tiles_l, ids_l, d_l = [], [], []
    for i in np.nonzero(live)[0]:
        ys, xs = np.meshgrid(np.arange(y0[i], y1[i]), np.arange(x0[i], x1[i]), indexing="ij")
        t = (ys * gx + xs).reshape(-1)
        tiles_l.append(t)
        ids_l.append(np.full(t.shape, i, dtype=np.int64))
        d_l.append(np.full(t.shape, dbits[i], dtype=np.uint64))
    if not tiles_l:
        return torch.zeros(0, dtype=torch.int32, device=uv.device), torch.zeros(T, 2, dtype=torch.int32, device=uv.device)
    tiles = np.concatenate(tiles_l).astype(np.uint64)
    ids = np.concatenate(ids_l)
    keys = (tiles << np.uint64(32)) | np.concatenate(d_l)
    order = np.argsort(keys, kind="stable")
    ids_sorted = ids[order].astype(np.int32)
    tiles_sorted = tiles[order].astype(np.int64)
    rng = np.zeros((T, 2), dtype=np.int32)
    starts = np.searchsorted(tiles_sorted, np.arange(T), side="left")
    ends = np.searchsorted(tiles_sorted, np.arange(T), side="right")
    nz = ends > starts
    rng[nz, 0] = starts[nz]
    rng[nz, 1] = ends[nz]
    return torch.from_numpy(ids_sorted).to(uv.device), torch.from_numpy(rng).to(uv.device)


# --------------------------------------------------------------------------- a6 / a7
def alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc=None,
                   return_aux=False):
    """Front-to-back compositing of C-channel features per pixel.

    Call sites: /root/reference/gflow/utils/render.py:58-64,68-74,84-90,99-105,148-154.
    alpha = min(0.99, o * exp(-0.5 d^T Q d)); skipped when the exponent is
    positive or alpha < 1/255; a pixel stops *before* applying the Gaussian that
    would push T below 1e-4; out_c = sum f_c alpha T + T_final * bg.  Pixel
    centres are at integer coordinates.  The 0.99 clamp is transparent to the
    backward pass (3DGS-lineage behaviour).  Output (C,H,W).
    ``ndc`` is accepted for signature compatibility and ignored.
    """
    C = feature.shape[1]
    dt = feature.dtype
    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    rows = []
    dev = feature.device  # the oracle is device agnostic: with CUDA tensors it doubles as the eager-PyTorch proxy
    finalT = torch.ones(H, W, dtype=dt, device=dev)
    ncontrib = torch.zeros(H, W, dtype=torch.int32, device=dev)
    rng = tile_range.to(torch.int64).cpu()
    ids_all = gaussian_ids_sorted.to(torch.int64)
    for ty in range(gy):
        row = []
        for tx in range(gx):
            t = ty * gx + tx
            s, e = int(rng[t, 0]), int(rng[t, 1])
            hh = min(TILE, H - ty * TILE)
            ww = min(TILE, W - tx * TILE)
            if e <= s:
                row.append(torch.full((C, hh, ww), float(bg), dtype=dt, device=dev))
                continue
            ids = ids_all[s:e]
            py, px = torch.meshgrid(
                torch.arange(ty * TILE, ty * TILE + hh, dtype=dt, device=dev),
                torch.arange(tx * TILE, tx * TILE + ww, dtype=dt, device=dev),
                indexing="ij",
            )
            px = px.reshape(-1, 1)
            py = py.reshape(-1, 1)
            g_uv = uv[ids]
            g_con = conic[ids]
            g_op = opacity[ids].reshape(1, -1)
            g_f = feature[ids]
            dx = g_uv[:, 0].reshape(1, -1) - px
            dy = g_uv[:, 1].reshape(1, -1) - py
            power = -0.5 * (g_con[:, 0] * dx * dx + g_con[:, 2] * dy * dy) - g_con[:, 1] * dx * dy
            raw = g_op * torch.exp(power)
            alpha = raw - (raw - ALPHA_MAX).clamp(min=0).detach()  # min(0.99, raw), straight-through
            valid = (power <= 0) & (alpha >= ALPHA_MIN)
            a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
            one_m = 1.0 - a_eff
            cum = torch.cumprod(one_m, dim=1)  # transmittance after j
            stop = valid & (cum.detach() < T_EPS)
            excluded = torch.cumsum(stop.to(torch.int32), dim=1) > 0
            a_inc = torch.where(excluded, torch.zeros_like(a_eff), a_eff)
            cum_inc = torch.cumprod(1.0 - a_inc, dim=1)
            t_before = torch.cat([torch.ones_like(cum_inc[:, :1]), cum_inc[:, :-1]], dim=1)
            wgt = a_inc * t_before
            t_fin = cum_inc[:, -1]
            out = wgt @ g_f + t_fin.unsqueeze(1) * float(bg)
            row.append(out.t().reshape(C, hh, ww))
            if return_aux:
                n = e - s
                contrib = (wgt.detach() > 0) | (valid & ~excluded)
                idx = torch.arange(1, n + 1, dtype=torch.int32, device=dev).reshape(1, -1)
                last = torch.where(contrib, idx, torch.zeros_like(idx)).max(dim=1).values
                finalT[ty * TILE: ty * TILE + hh, tx * TILE: tx * TILE + ww] = t_fin.detach().reshape(hh, ww)
                ncontrib[ty * TILE: ty * TILE + hh, tx * TILE: tx * TILE + ww] = last.reshape(hh, ww)
        rows.append(torch.cat(row, dim=2))
    img = torch.cat(rows, dim=1)
    if return_aux:
        return img, finalT, ncontrib
    return img


# --------------------------------------------------------------------------- a8
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def sh_basis(dirs, n_coef):
    """Real SH basis (3DGS constants) evaluated at unit directions; (N, n_coef)."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    out = [torch.full_like(x, SH_C0)]
    if n_coef > 1:
        out += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if n_coef > 4:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if n_coef > 9:
        out += [
            SH_C3[0] * y * (3.0 * xx - yy),
            SH_C3[1] * xy * z,
            SH_C3[2] * y * (4.0 * zz - xx - yy),
            SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy),
            SH_C3[4] * x * (4.0 * zz - xx - yy),
            SH_C3[5] * z * (xx - yy),
            SH_C3[6] * x * (xx - 3.0 * yy),
        ]
    return torch.stack(out, dim=1)


def compute_sh(shs, dirs, visible=None):
    """colour_c = sum_k shs[:, c, k] * Y_k(dir / |dir|); degree from K in {1,4,9,16}.

    Not called by GFlow (colour is sigmoid(rgb), trainer.py:68); part of the
    north_star operator surface and BASELINE config 5.  shs (N,C,K), dirs (N,3)
    un-normalised view directions; no +0.5 / clamp inside.
    """
    n_coef = shs.shape[2]
    assert n_coef in (1, 4, 9, 16)
    d = dirs / torch.sqrt((dirs * dirs).sum(dim=1, keepdim=True))
    Y = sh_basis(d, n_coef)
    out = (shs * Y.unsqueeze(1)).sum(dim=2)
    if visible is not None:
        out = torch.where(visible.reshape(-1, 1), out, torch.zeros_like(out))
    return out


# --------------------------------------------------------------------------- chain
def render_step(xyz, scale, rotate, opacity, feature, intr, extr, bg, W, H):
    """The op chain of /root/reference/gflow/utils/render.py:21-64 ('rgb' only)."""
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    ids, rng = sort_gaussian(uv, depth, W, H, radius, tiles)
    img = alpha_blending(uv, conic, opacity, feature, ids, rng, bg, W, H)
    return img, dict(uv=uv, depth=depth, cov3d=cov3d, conic=conic, radius=radius, tiles=tiles, ids=ids, rng=rng)
