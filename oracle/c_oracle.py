"""ctypes front end of the plain-C oracle (``oracle/splat_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of splat_oracle.c.  PARITY UNPINNED
(no msplat source, tests or golden vectors exist under /root/reference).

All functions take / return CPU torch tensors (float32 / int32 / bool) with the
shapes of the msplat operator surface.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsplat_oracle.so")
_lib = None

c_f = ctypes.c_float
c_i = ctypes.c_int
c_p = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_sort_gaussian.restype = ctypes.c_int64
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _f(t):
    return t.detach().to(torch.float32).contiguous()


def _p(t):
    return c_p(t.data_ptr()) if t is not None else c_p(0)


def _vis(visible):
    if visible is None:
        return None
    return visible.reshape(-1).to(torch.uint8).contiguous()


def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):
    xyz, intr, extr = _f(xyz), _f(intr), _f(extr)
    N = xyz.shape[0]
    uv = torch.empty(N, 2)
    depth = torch.empty(N, 1)
    lib().orc_project_point_fwd(_p(xyz), _p(intr), _p(extr), c_i(N), c_i(W), c_i(H), c_f(nearest), c_f(extent),
                                _p(uv), _p(depth))
    return uv, depth


def project_point_bwd(xyz, intr, extr, W, H, g_uv, g_depth, nearest=0.2, extent=1.3):
    xyz, intr, extr, g_uv = _f(xyz), _f(intr), _f(extr), _f(g_uv)
    g_depth = _f(g_depth) if g_depth is not None else None
    N = xyz.shape[0]
    d_xyz, d_intr, d_extr = torch.empty(N, 3), torch.empty(4), torch.empty(3, 4)
    lib().orc_project_point_bwd(_p(xyz), _p(intr), _p(extr), c_i(N), c_i(W), c_i(H), c_f(nearest), c_f(extent),
                                _p(g_uv), _p(g_depth), _p(d_xyz), _p(d_intr), _p(d_extr))
    return d_xyz, d_intr, d_extr


def compute_cov3d(scale, rotate, visible=None):
    scale, rotate, vis = _f(scale), _f(rotate), _vis(visible)
    N = scale.shape[0]
    cov = torch.empty(N, 6)
    lib().orc_compute_cov3d_fwd(_p(scale), _p(rotate), _p(vis), c_i(N), _p(cov))
    return cov


def compute_cov3d_bwd(scale, rotate, visible, g_cov):
    scale, rotate, vis, g_cov = _f(scale), _f(rotate), _vis(visible), _f(g_cov)
    N = scale.shape[0]
    d_s, d_q = torch.empty(N, 3), torch.empty(N, 4)
    lib().orc_compute_cov3d_bwd(_p(scale), _p(rotate), _p(vis), c_i(N), _p(g_cov), _p(d_s), _p(d_q))
    return d_s, d_q


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):
    xyz, cov3d, intr, extr, uv, vis = _f(xyz), _f(cov3d), _f(intr), _f(extr), _f(uv), _vis(visible)
    N = xyz.shape[0]
    conic = torch.empty(N, 3)
    radius = torch.empty(N, 1, dtype=torch.int32)
    tiles = torch.empty(N, 1, dtype=torch.int32)
    lib().orc_ewa_project_fwd(_p(xyz), _p(cov3d), _p(intr), _p(extr), _p(uv), c_i(N), c_i(W), c_i(H), _p(vis),
                              _p(conic), _p(radius), _p(tiles))
    return conic, radius, tiles


def ewa_project_bwd(xyz, cov3d, intr, extr, uv, W, H, visible, g_conic):
    xyz, cov3d, intr, extr, uv, vis = _f(xyz), _f(cov3d), _f(intr), _f(extr), _f(uv), _vis(visible)
    g_conic = _f(g_conic)
    N = xyz.shape[0]
    d_xyz, d_cov, d_intr, d_extr = torch.empty(N, 3), torch.empty(N, 6), torch.empty(4), torch.empty(3, 4)
    lib().orc_ewa_project_bwd(_p(xyz), _p(cov3d), _p(intr), _p(extr), _p(uv), c_i(N), c_i(W), c_i(H), _p(vis),
                              _p(g_conic), _p(d_xyz), _p(d_cov), _p(d_intr), _p(d_extr))
    return d_xyz, d_cov, d_intr, d_extr


def sort_gaussian(uv, depth, W, H, radius, tiles_touched):
    uv, depth = _f(uv), _f(depth)
    radius = radius.reshape(-1).to(torch.int32).contiguous()
    tiles = tiles_touched.reshape(-1).to(torch.int32).contiguous()
    N = uv.shape[0]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    K = int(lib().orc_sort_gaussian(_p(uv), _p(depth), c_i(N), c_i(W), c_i(H), _p(radius), _p(tiles), c_p(0), c_p(0)))
    ids = torch.empty(K, dtype=torch.int32)
    rng = torch.zeros(T, 2, dtype=torch.int32)
    K2 = int(lib().orc_sort_gaussian(_p(uv), _p(depth), c_i(N), c_i(W), c_i(H), _p(radius), _p(tiles), _p(ids), _p(rng)))
    return ids[:K2], rng


def alpha_blending(uv, conic, opacity, feature, ids, rng, bg, W, H, ndc=None, return_aux=False):
    uv, conic, opacity, feature = _f(uv), _f(conic), _f(opacity), _f(feature)
    ids, rng = ids.to(torch.int32).contiguous(), rng.to(torch.int32).contiguous()
    C = feature.shape[1]
    out = torch.empty(C, H, W)
    fT = torch.empty(H, W)
    nc = torch.empty(H, W, dtype=torch.int32)
    lib().orc_alpha_blending_fwd(_p(uv), _p(conic), _p(opacity), _p(feature), c_i(C), _p(ids), _p(rng), c_f(bg),
                                 c_i(W), c_i(H), _p(out), _p(fT), _p(nc))
    if return_aux:
        return out, fT, nc
    return out


def alpha_blending_bwd(uv, conic, opacity, feature, ids, rng, bg, W, H, final_T, n_contrib, g_out):
    uv, conic, opacity, feature = _f(uv), _f(conic), _f(opacity), _f(feature)
    ids, rng = ids.to(torch.int32).contiguous(), rng.to(torch.int32).contiguous()
    final_T, n_contrib, g_out = _f(final_T), n_contrib.to(torch.int32).contiguous(), _f(g_out)
    N, C = feature.shape
    d_uv, d_conic, d_op, d_f = torch.empty(N, 2), torch.empty(N, 3), torch.empty(N, 1), torch.empty(N, C)
    lib().orc_alpha_blending_bwd(_p(uv), _p(conic), _p(opacity), _p(feature), c_i(C), c_i(N), _p(ids), _p(rng),
                                 c_f(bg), c_i(W), c_i(H), _p(final_T), _p(n_contrib), _p(g_out), _p(d_uv),
                                 _p(d_conic), _p(d_op), _p(d_f))
    return d_uv, d_conic, d_op, d_f


def compute_sh(shs, dirs, visible=None):
    shs, dirs, vis = _f(shs), _f(dirs), _vis(visible)
    N, C, K = shs.shape
    out = torch.empty(N, C)
    lib().orc_compute_sh_fwd(_p(shs), _p(dirs), _p(vis), c_i(N), c_i(C), c_i(K), _p(out))
    return out


def compute_sh_bwd(shs, dirs, visible, g_out):
    shs, dirs, vis, g_out = _f(shs), _f(dirs), _vis(visible), _f(g_out)
    N, C, K = shs.shape
    d_shs, d_dirs = torch.empty(N, C, K), torch.empty(N, 3)
    lib().orc_compute_sh_bwd(_p(shs), _p(dirs), _p(vis), c_i(N), c_i(C), c_i(K), _p(g_out), _p(d_shs), _p(d_dirs))
    return d_shs, d_dirs


def render_step_fwd_bwd(xyz, scale, rotate, opacity, feature, intr, extr, bg, W, H, g_img):
    """One render step (SURVEY.md 8d unit (ii)) forward + backward on the CPU.

    Returns (img, grads dict).  Used as the CPU baseline of bench.py.
    """
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    ids, rng = sort_gaussian(uv, depth, W, H, radius, tiles)
    img, fT, nc = alpha_blending(uv, conic, opacity, feature, ids, rng, bg, W, H, return_aux=True)
    d_uv, d_conic, d_op, d_f = alpha_blending_bwd(uv, conic, opacity, feature, ids, rng, bg, W, H, fT, nc, g_img)
    d_xyz2, d_cov, d_intr2, d_extr2 = ewa_project_bwd(xyz, cov3d, intr, extr, uv, W, H, visible, d_conic)
    d_s, d_q = compute_cov3d_bwd(scale, rotate, visible, d_cov)
    d_xyz1, d_intr1, d_extr1 = project_point_bwd(xyz, intr, extr, W, H, d_uv, None)
    grads = dict(xyz=d_xyz1 + d_xyz2, scale=d_s, rotate=d_q, opacity=d_op, feature=d_f,
                 intr=d_intr1 + d_intr2, extr=d_extr1 + d_extr2)
    return img, grads, dict(K=int(ids.numel()))
