"""Host-side mirror of the `msplat` operator surface, backed by libgflow_b200.so.

Same names, positional arguments, return tuples, dtypes and error behaviour as the
operators GFlow calls (/root/reference/gflow/utils/render.py:21-154,
/root/reference/gflow/trainer.py:955):

    project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3) -> (uv, depth)
    compute_cov3d(scale, rotate, visible=None)                    -> cov3d
    ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None)   -> (conic, radius, tiles_touched)
    sort_gaussian(uv, depth, W, H, radius, tiles_touched)         -> (gaussian_ids_sorted, tile_range)
    alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc=None)
                                                                  -> feature_map (C,H,W)
    compute_sh(shs, dirs, visible=None)                           -> (N,C)
    rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg) -> (C,H,W)

Each differentiable op is a torch.autograd.Function whose forward / backward call the C ABI
with raw device pointers on the current CUDA stream.  PyTorch is plumbing only (device
memory, streams, autograd graph).  There is no CPU path: CPU tensors raise RuntimeError.
"""
from __future__ import annotations

import threading

import torch

from . import capi

_lib = capi.load()
TILE = 16


# --------------------------------------------------------------------------- helpers
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device, as an int."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs ~10 us)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _prep(t: torch.Tensor, name: str, dtype=torch.float32, shape=None) -> torch.Tensor:
    """Validate device / dtype / shape and return a contiguous, 16-byte aligned tensor."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"gflow_b200: {name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"gflow_b200: {name} must be a CUDA tensor (no CPU fallback exists), got device {t.device}")
    if t.dtype != dtype:
        raise RuntimeError(f"gflow_b200: {name} must have dtype {dtype}, got {t.dtype}")
    if shape is not None:
        if t.dim() != len(shape) or any(s is not None and int(d) != s for d, s in zip(t.shape, shape)):
            raise RuntimeError(f"gflow_b200: {name} must have shape {shape}, got {tuple(t.shape)}")
    if t.requires_grad:
        t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _vis(visible, N, device):
    if visible is None:
        return None
    if not visible.is_cuda:
        raise RuntimeError("gflow_b200: visible must be a CUDA tensor")
    if visible.numel() != N:
        raise RuntimeError(f"gflow_b200: visible must have {N} elements, got {visible.numel()}")
    v = visible.detach().reshape(-1)
    if v.dtype == torch.bool:
        v = v.contiguous().view(torch.uint8)
    else:
        v = (v != 0).view(torch.uint8)
    return v


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _same_device(*ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"gflow_b200: tensors on different devices ({dev} vs {t.device})")
    return dev


def _grid(W: int, H: int):
    return (W + TILE - 1) // TILE, (H + TILE - 1) // TILE


# --------------------------------------------------------------------------- project_point
class _ProjectPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, intr, extr, W, H, nearest, extent):
        xyz_c = _prep(xyz, "xyz", shape=(None, 3))
        intr_c = _prep(intr, "intr", shape=(4,))
        extr_c = _prep(extr, "extr", shape=(3, 4))
        dev = _same_device(xyz_c, intr_c, extr_c)
        N = xyz_c.shape[0]
        with _on_device(dev):
            uv = torch.empty(N, 2, device=dev, dtype=torch.float32)
            depth = torch.empty(N, 1, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_project_point_fwd(xyz_c.data_ptr(), intr_c.data_ptr(), extr_c.data_ptr(), N, W, H,
                                                  nearest, extent, uv.data_ptr(), depth.data_ptr(), _stream()),
                       "project_point forward")
        ctx.save_for_backward(xyz_c, intr_c, extr_c)
        ctx.meta = (W, H, nearest, extent)
        return uv, depth

    @staticmethod
    def backward(ctx, g_uv, g_depth):
        xyz, intr, extr = ctx.saved_tensors
        W, H, nearest, extent = ctx.meta
        N = xyz.shape[0]
        dev = xyz.device
        with _on_device(dev):
            g_uv = torch.zeros(N, 2, device=dev) if g_uv is None else _prep(g_uv, "grad uv")
            g_depth = None if g_depth is None else _prep(g_depth, "grad depth")
            d_xyz = torch.empty(N, 3, device=dev, dtype=torch.float32)
            d_cam = torch.empty(16, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_project_point_bwd(xyz.data_ptr(), intr.data_ptr(), extr.data_ptr(), N, W, H, nearest,
                                                  extent, g_uv.data_ptr(), _ptr(g_depth), d_xyz.data_ptr(),
                                                  d_cam.data_ptr(), _stream()), "project_point backward")
        return d_xyz, d_cam[12:16], d_cam[:12].view(3, 4), None, None, None, None


def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):
    """msplat.project_point -- /root/reference/gflow/utils/render.py:21-24, trainer.py:955."""
    return _ProjectPoint.apply(xyz, intr, extr, int(W), int(H), float(nearest), float(extent))


# --------------------------------------------------------------------------- compute_cov3d
class _ComputeCov3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale, rotate, visible):
        scale_c = _prep(scale, "scale", shape=(None, 3))
        rotate_c = _prep(rotate, "rotate", shape=(None, 4))
        dev = _same_device(scale_c, rotate_c)
        N = scale_c.shape[0]
        if rotate_c.shape[0] != N:
            raise RuntimeError("gflow_b200: scale and rotate must have the same number of rows")
        vis = _vis(visible, N, dev)
        with _on_device(dev):
            cov3d = torch.empty(N, 6, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_compute_cov3d_fwd(scale_c.data_ptr(), rotate_c.data_ptr(), _ptr(vis), N,
                                                  cov3d.data_ptr(), _stream()), "compute_cov3d forward")
        ctx.save_for_backward(scale_c, rotate_c)
        ctx.vis = vis
        return cov3d

    @staticmethod
    def backward(ctx, g_cov):
        scale, rotate = ctx.saved_tensors
        N = scale.shape[0]
        dev = scale.device
        with _on_device(dev):
            g_cov = _prep(g_cov, "grad cov3d")
            d_scale = torch.empty(N, 3, device=dev, dtype=torch.float32)
            d_rotate = torch.empty(N, 4, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_compute_cov3d_bwd(scale.data_ptr(), rotate.data_ptr(), _ptr(ctx.vis), N,
                                                  g_cov.data_ptr(), d_scale.data_ptr(), d_rotate.data_ptr(),
                                                  _stream()), "compute_cov3d backward")
        return d_scale, d_rotate, None


def compute_cov3d(scale, rotate, visible=None):
    """msplat.compute_cov3d -- /root/reference/gflow/utils/render.py:37-41."""
    return _ComputeCov3D.apply(scale, rotate, visible)


# --------------------------------------------------------------------------- ewa_project
class _EwaProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, cov3d, intr, extr, uv, W, H, visible):
        xyz_c = _prep(xyz, "xyz", shape=(None, 3))
        N = xyz_c.shape[0]
        cov_c = _prep(cov3d, "cov3d", shape=(N, 6))
        intr_c = _prep(intr, "intr", shape=(4,))
        extr_c = _prep(extr, "extr", shape=(3, 4))
        uv_c = _prep(uv, "uv", shape=(N, 2))
        dev = _same_device(xyz_c, cov_c, intr_c, extr_c, uv_c)
        vis = _vis(visible, N, dev)
        with _on_device(dev):
            conic = torch.empty(N, 3, device=dev, dtype=torch.float32)
            radius = torch.empty(N, 1, device=dev, dtype=torch.int32)
            tiles = torch.empty(N, 1, device=dev, dtype=torch.int32)
            capi.check(_lib.gfb_ewa_project_fwd(xyz_c.data_ptr(), cov_c.data_ptr(), intr_c.data_ptr(),
                                                extr_c.data_ptr(), uv_c.data_ptr(), N, W, H, _ptr(vis),
                                                conic.data_ptr(), radius.data_ptr(), tiles.data_ptr(), _stream()),
                       "ewa_project forward")
        ctx.save_for_backward(xyz_c, cov_c, intr_c, extr_c, uv_c)
        ctx.vis = vis
        ctx.meta = (W, H)
        ctx.mark_non_differentiable(radius, tiles)
        return conic, radius, tiles

    @staticmethod
    def backward(ctx, g_conic, _g_radius, _g_tiles):
        xyz, cov3d, intr, extr, uv = ctx.saved_tensors
        W, H = ctx.meta
        N = xyz.shape[0]
        dev = xyz.device
        with _on_device(dev):
            g_conic = _prep(g_conic, "grad conic")
            d_xyz = torch.empty(N, 3, device=dev, dtype=torch.float32)
            d_cov = torch.empty(N, 6, device=dev, dtype=torch.float32)
            d_cam = torch.empty(16, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_ewa_project_bwd(xyz.data_ptr(), cov3d.data_ptr(), intr.data_ptr(), extr.data_ptr(),
                                                uv.data_ptr(), N, W, H, _ptr(ctx.vis), g_conic.data_ptr(),
                                                d_xyz.data_ptr(), d_cov.data_ptr(), d_cam.data_ptr(), _stream()),
                       "ewa_project backward")
        return d_xyz, d_cov, d_cam[12:16], d_cam[:12].view(3, 4), None, None, None, None


def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):
    """msplat.ewa_project -- /root/reference/gflow/utils/render.py:44-49."""
    return _EwaProject.apply(xyz, cov3d, intr, extr, uv, int(W), int(H), visible)


# --------------------------------------------------------------------------- sort_gaussian
import ctypes as _ctypes

_K_HINT = {}  # (device index, N, W, H) -> last K: sizes the speculative buffers of the next call
GFB_E_CAPACITY = -3


def _capacity_for(key, N):
    k = _K_HINT.get(key)
    return (4 * N + 4096) if k is None else (k + k // 4 + 4096)


@torch.no_grad()
def sort_gaussian(uv, depth, W, H, radius, tiles_touched):
    """msplat.sort_gaussian -- /root/reference/gflow/utils/render.py:52-54.

    Returns gaussian_ids_sorted (K,) int32 and tile_range (T,2) int32.  K must reach the host because
    the API returns a tensor of exactly K entries; the read-back is hidden behind the scatter + sort
    kernels, which are enqueued speculatively with the previous call's K (plus slack) as capacity.
    """
    W, H = int(W), int(H)
    uv_c = _prep(uv, "uv", shape=(None, 2))
    N = uv_c.shape[0]
    depth_c = _prep(depth, "depth").reshape(-1)
    radius_c = _prep(radius, "radius", dtype=torch.int32).reshape(-1)
    tiles_c = _prep(tiles_touched, "tiles_touched", dtype=torch.int32).reshape(-1)
    if depth_c.numel() != N or radius_c.numel() != N or tiles_c.numel() != N:
        raise RuntimeError("gflow_b200: uv, depth, radius and tiles_touched must describe the same N Gaussians")
    dev = _same_device(uv_c, depth_c, radius_c, tiles_c)
    gx, gy = _grid(W, H)
    T = gx * gy
    key = (dev.index, N, W, H)
    cap = _capacity_for(key, N)
    k_host = _ctypes.c_int64(0)
    with _on_device(dev):
        tile_ws = torch.empty(_lib.gfb_sort_tile_workspace_bytes(W, H), device=dev, dtype=torch.uint8)
        tile_range = torch.empty(T, 2, device=dev, dtype=torch.int32)
        while True:
            keys = torch.empty(max(cap, 1), device=dev, dtype=torch.int64)
            ids = torch.empty(max(cap, 1), device=dev, dtype=torch.int32)
            rc = _lib.gfb_sort_gaussian(uv_c.data_ptr(), depth_c.data_ptr(), radius_c.data_ptr(), tiles_c.data_ptr(),
                                        N, W, H, tile_ws.data_ptr(), cap, keys.data_ptr(), ids.data_ptr(),
                                        tile_range.data_ptr(),
                                        _ctypes.byref(k_host), _stream())
            K = int(k_host.value)
            if rc == GFB_E_CAPACITY:
                cap = K + K // 8 + 1024
                continue
            capi.check(rc, "sort_gaussian")
            break
    _K_HINT[key] = K
    return ids[:K], tile_range


# --------------------------------------------------------------------------- alpha_blending
class _GeomStreamCache:
    """Last packed geometry stream, keyed on tensor identity + version counters.

    render_multiple blends rgb, depth and depth-colour with the same (uv, conic, opacity,
    gaussian_ids_sorted) objects (/root/reference/gflow/utils/render.py:58-90); the packed
    stream is built once and shared.  Strong references pin the key tensors so an address can
    never be recycled under a stale entry.
    """

    def __init__(self):
        self.key = None
        self.refs = None
        self.stream = None

    @staticmethod
    def _key(uv, conic, opacity, ids):
        # the CUDA stream is part of the key: a stream packed on one stream must not be read from another unsynchronised
        return (id(uv), uv._version, id(conic), conic._version, id(opacity), opacity._version, id(ids), ids._version,
                _stream() if uv.is_cuda else 0)

    def get(self, uv, conic, opacity, ids):
        if self.key == self._key(uv, conic, opacity, ids):
            return self.stream
        return None

    def put(self, uv, conic, opacity, ids, stream):
        self.key = self._key(uv, conic, opacity, ids)
        self.refs = (uv, conic, opacity, ids)
        self.stream = stream

    def clear(self):
        self.key = self.refs = self.stream = None


_geom_cache = _GeomStreamCache()


def _pack_geometry(uv_c, conic_c, opacity_c, ids_c, K):
    dev = uv_c.device
    stream = torch.empty(max(K, 1) * 8, device=dev, dtype=torch.float32)
    capi.check(_lib.gfb_blend_pack_geometry(uv_c.data_ptr(), conic_c.data_ptr(), opacity_c.data_ptr(),
                                            ids_c.data_ptr(), K, stream.data_ptr(), _stream()), "blend pack geometry")
    return stream


class _AlphaBlending(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, conic, opacity, feature, ids, tile_range, bg, W, H, ndc, geom_stream):
        uv_c = _prep(uv, "uv", shape=(None, 2))
        N = uv_c.shape[0]
        conic_c = _prep(conic, "conic", shape=(N, 3))
        opacity_c = _prep(opacity, "opacity").reshape(-1)
        if opacity_c.numel() != N:
            raise RuntimeError(f"gflow_b200: opacity must have {N} elements, got {opacity_c.numel()}")
        feature_c = _prep(feature, "feature", shape=(N, None))
        C = feature_c.shape[1]
        if C < 1:
            raise RuntimeError("gflow_b200: feature needs at least one channel")
        ids_c = _prep(ids, "gaussian_ids_sorted", dtype=torch.int32).reshape(-1)
        gx, gy = _grid(W, H)
        T = gx * gy
        tr_c = _prep(tile_range, "tile_range", dtype=torch.int32, shape=(T, 2))
        dev = _same_device(uv_c, conic_c, opacity_c, feature_c, ids_c, tr_c)
        K = ids_c.numel()
        st = _stream
        with _on_device(dev):
            if geom_stream is None:
                geom_stream = _pack_geometry(uv_c, conic_c, opacity_c, ids_c, K)
            out = torch.empty(C, H, W, device=dev, dtype=torch.float32)
            final_T = torch.empty(H, W, device=dev, dtype=torch.float32)
            n_contrib = torch.empty(H, W, device=dev, dtype=torch.int32)
            feat_streams = []
            for c0 in range(0, C, 4):
                cg = min(4, C - c0)
                fs = torch.empty(max(K, 1) * 4, device=dev, dtype=torch.float32)
                capi.check(_lib.gfb_blend_pack_feature(feature_c.data_ptr(), C, c0, cg, ids_c.data_ptr(), K,
                                                       fs.data_ptr(), st()), "blend pack feature")
                capi.check(_lib.gfb_alpha_blending_fwd(geom_stream.data_ptr(), fs.data_ptr(), K, tr_c.data_ptr(), C,
                                                       c0, cg, bg, W, H, out.data_ptr(), final_T.data_ptr(),
                                                       n_contrib.data_ptr(), st()), "alpha_blending forward")
                feat_streams.append(fs)
        ctx.save_for_backward(geom_stream, ids_c, tr_c, final_T, n_contrib, *feat_streams)
        ctx.meta = (N, C, K, float(bg), W, H)
        ctx.ndc_grad = ndc is not None and isinstance(ndc, torch.Tensor) and ndc.requires_grad
        return out

    @staticmethod
    def backward(ctx, g_out):
        geom_stream, ids_c, tr_c, final_T, n_contrib, *feat_streams = ctx.saved_tensors
        N, C, K, bg, W, H = ctx.meta
        dev = geom_stream.device
        st = _stream
        with _on_device(dev):
            g_out = _prep(g_out, "grad feature_map", shape=(C, H, W))
            d_uv = torch.empty(N, 2, device=dev, dtype=torch.float32)
            d_conic = torch.empty(N, 3, device=dev, dtype=torch.float32)
            d_opacity = torch.empty(N, 1, device=dev, dtype=torch.float32)
            d_feature = torch.empty(N, C, device=dev, dtype=torch.float32)
            for gi, c0 in enumerate(range(0, C, 4)):
                cg = min(4, C - c0)
                grad_pack = torch.zeros(max(N, 1) * 12, device=dev, dtype=torch.float32)
                capi.check(_lib.gfb_alpha_blending_bwd(geom_stream.data_ptr(), feat_streams[gi].data_ptr(), K,
                                                       ids_c.data_ptr(), tr_c.data_ptr(), C, c0, cg, bg, W, H,
                                                       final_T.data_ptr(), n_contrib.data_ptr(), g_out.data_ptr(),
                                                       grad_pack.data_ptr(), st()), "alpha_blending backward")
                capi.check(_lib.gfb_blend_unpack_grads(grad_pack.data_ptr(), N, C, c0, cg, d_uv.data_ptr(),
                                                       d_conic.data_ptr(), d_opacity.data_ptr(), d_feature.data_ptr(),
                                                       1 if gi > 0 else 0, st()), "alpha_blending unpack grads")
        d_ndc = d_uv.clone() if ctx.ndc_grad else None
        return d_uv, d_conic, d_opacity, d_feature, None, None, None, None, None, d_ndc, None


def alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc=None):
    """msplat.alpha_blending -- /root/reference/gflow/utils/render.py:58-64 (and 68-105, 148-154).

    ``bg`` is a Python float; the result is (C,H,W).  ``ndc`` (optional, upstream's hook for
    densification statistics) receives a copy of the screen-space gradient of ``uv``.
    """
    W, H = int(W), int(H)
    geom_stream = None
    if all(isinstance(t, torch.Tensor) and t.is_cuda for t in (uv, conic, opacity, gaussian_ids_sorted)):
        geom_stream = _geom_cache.get(uv, conic, opacity, gaussian_ids_sorted)
        if geom_stream is None and uv.dtype == conic.dtype == opacity.dtype == torch.float32 \
                and gaussian_ids_sorted.dtype == torch.int32 and uv.dim() == 2 and conic.dim() == 2 \
                and conic.shape == (uv.shape[0], 3) and opacity.numel() == uv.shape[0]:
            with torch.no_grad(), _on_device(uv.device):
                geom_stream = _pack_geometry(_prep(uv, "uv", shape=(None, 2)), _prep(conic, "conic"),
                                             _prep(opacity, "opacity").reshape(-1),
                                             _prep(gaussian_ids_sorted, "gaussian_ids_sorted", dtype=torch.int32)
                                             .reshape(-1), gaussian_ids_sorted.numel())
            _geom_cache.put(uv, conic, opacity, gaussian_ids_sorted, geom_stream)
    return _AlphaBlending.apply(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, float(bg), W, H, ndc,
                                geom_stream)


# --------------------------------------------------------------------------- compute_sh
class _ComputeSH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shs, dirs, visible):
        shs_c = _prep(shs, "shs", shape=(None, None, None))
        N, C, K = shs_c.shape
        if K not in (1, 4, 9, 16):
            raise RuntimeError(f"gflow_b200: shs last dimension must be 1, 4, 9 or 16 (degree 0..3), got {K}")
        dirs_c = _prep(dirs, "dirs", shape=(N, 3))
        dev = _same_device(shs_c, dirs_c)
        vis = _vis(visible, N, dev)
        with _on_device(dev):
            out = torch.empty(N, C, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_compute_sh_fwd(shs_c.data_ptr(), dirs_c.data_ptr(), _ptr(vis), N, C, K,
                                               out.data_ptr(), _stream()), "compute_sh forward")
        ctx.save_for_backward(shs_c, dirs_c)
        ctx.vis = vis
        return out

    @staticmethod
    def backward(ctx, g_out):
        shs, dirs = ctx.saved_tensors
        N, C, K = shs.shape
        dev = shs.device
        with _on_device(dev):
            g_out = _prep(g_out, "grad sh colour", shape=(N, C))
            d_shs = torch.empty(N, C, K, device=dev, dtype=torch.float32)
            d_dirs = torch.empty(N, 3, device=dev, dtype=torch.float32)
            capi.check(_lib.gfb_compute_sh_bwd(shs.data_ptr(), dirs.data_ptr(), _ptr(ctx.vis), N, C, K,
                                               g_out.data_ptr(), d_shs.data_ptr(), d_dirs.data_ptr(), _stream()),
                       "compute_sh backward")
        return d_shs, d_dirs, None


def compute_sh(shs, dirs, visible=None):
    """msplat.compute_sh (north_star surface; not called by GFlow): shs (N,C,K), dirs (N,3)."""
    return _ComputeSH.apply(shs, dirs, visible)


# --------------------------------------------------------------------------- rasterization
import os as _os
import warnings as _warnings

# Lazy validation of the speculative K (GFLOW_B200_LAZY_K=0 switches it off): in a training loop the forward
# does not wait for `preprocess` to deliver K -- the capacity guessed from the previous call (+25 %) is checked
# when the backward starts, by which time the K ticket has long completed.  The host then never blocks on the
# GPU inside a step and can run ahead of it.  If K did outgrow the guess (it would have to grow by a quarter
# between two consecutive calls), the backward re-runs the forward with the right capacity before computing
# gradients and warns that the image already handed out missed the tail of some tile lists.
# It is only used for a call that continues a loop over the SAME parameter tensor (the storage of `xyz` was seen by one
# of the last few calls of this size): a different scene of the same size takes the synchronous path, so an image is
# never clipped because an unrelated earlier call left a small K behind.
_LAZY_K = _os.environ.get("GFLOW_B200_LAZY_K", "1") != "0"
_LAZY_SEEN = {}  # (device index, N, W, H) -> data pointers of the last few xyz tensors rasterised at this size


def _continues_a_loop(key, xyz_c) -> bool:
    seen = _LAZY_SEEN.setdefault(key, [])
    ptr = xyz_c.data_ptr()
    hit = ptr in seen
    if not hit:
        seen.append(ptr)
        del seen[:-4]
    return hit
_CLIPPED_WARNING = ("gflow_b200.rasterization: the intersection count grew by more than 25 % between two consecutive "
                    "calls; the image returned by the earlier forward missed the tail of some tile lists (gradients were "
                    "recomputed from a corrected pass).  Set GFLOW_B200_LAZY_K=0 to validate K inside every forward.")


# Self-cleaning workspaces kept alive per (device, stream): the control block of the forward and the gradient pack of
# the backward return to all-zero by themselves (gfb_render_forward_keep / gfb_render_backward_keep), so a render step
# launches no memset.  A failed call drops them (they may be dirty).
_KEPT = {}
# ctypes releases the GIL inside a call: the kernels of one call on a kept block must reach the stream as one run
_KEEP_LOCK = threading.Lock()


def _kept(kind, dev, stream, size_key, nbytes):
    key = (kind, dev.index, stream, size_key)
    t = _KEPT.get(key)
    if t is None:
        t = torch.zeros(nbytes, device=dev, dtype=torch.uint8)
        _KEPT[key] = t
    return key, t


def _raster_forward(xyz_c, scale_c, rotate_c, opacity_c, feature_c, intr_c, extr_c, N, C, W, H, bg, nearest, extent, dev,
                    cap, out, lazy):
    """Enqueues the fused forward with capacity `cap`.  Returns (kbuf, tbuf, aux, cap, K, ticket): K is None and
    ticket names the pending hand-off when `lazy`, otherwise K is final (the call retried until it fitted)."""
    gx, gy = _grid(W, H)
    T = gx * gy
    k_host = _ctypes.c_int64(0)
    st = _stream()
    # per-Gaussian buffer (bytes): uv 8N | rect 8N | depth 4N | conic 12N | radius 4N
    gbuf = torch.empty(9 * max(N, 1), device=dev, dtype=torch.float32)
    gp = gbuf.data_ptr()
    p_uv, p_rect, p_depth, p_conic, p_radius = gp, gp + 8 * N, gp + 16 * N, gp + 20 * N, gp + 32 * N
    tbuf = torch.empty(8 * T, device=dev, dtype=torch.uint8)  # tile_range (T,2) int32: the backward reads it
    ckey, ctl = _kept("control", dev, st, (W, H), _lib.gfb_render_control_bytes(W, H))
    aux = torch.empty(2, H, W, device=dev, dtype=torch.float32)  # final_T | n_contrib (int32 bits)
    while True:
        # K-sized buffer (bytes): geom 32c | feat 16c | keys 8c | ids 4c
        kbuf = torch.empty(15 * max(cap, 1), device=dev, dtype=torch.float32)
        kp = kbuf.data_ptr()
        with _KEEP_LOCK:
            rc = _lib.gfb_render_forward_keep(
                xyz_c.data_ptr(), scale_c.data_ptr(), rotate_c.data_ptr(), opacity_c.data_ptr(),
                feature_c.data_ptr(), C, intr_c.data_ptr(), extr_c.data_ptr(), N, W, H, bg, nearest, extent,
                p_uv, p_depth, p_conic, p_radius, p_rect, ctl.data_ptr(), tbuf.data_ptr(), cap, kp + 48 * cap, kp + 56 * cap,
                kp, kp + 32 * cap, out.data_ptr(), aux.data_ptr(), aux.data_ptr() + 4 * H * W,
                None, st)
            if rc != 0:
                _KEPT.pop(ckey, None)
        capi.check(rc, "rasterization forward")
        ticket = int(_lib.gfb_k_ticket())
        if lazy:
            return kbuf, tbuf, aux, cap, None, ticket
        # everything is enqueued; pick up K (stored by the kernel into mapped pinned memory)
        capi.check(_lib.gfb_wait_k_ticket(ticket, _ctypes.byref(k_host)), "rasterization forward (K)")
        K = int(k_host.value)
        if K > cap:
            cap = K + K // 8 + 1024
            continue
        return kbuf, tbuf, aux, cap, K, ticket


class _Rasterize(torch.autograd.Function):
    """Fused chain (gfb_render_forward / gfb_render_backward): 4 + 2 kernels, no mid-pipeline drain."""

    @staticmethod
    def forward(ctx, xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, nearest, extent, lazy=False):
        xyz_c = _prep(xyz, "xyz", shape=(None, 3))
        N = xyz_c.shape[0]
        scale_c = _prep(scale, "scale", shape=(N, 3))
        rotate_c = _prep(rotate, "rotate", shape=(N, 4))
        opacity_c = _prep(opacity, "opacity").reshape(-1)
        if opacity_c.numel() != N:
            raise RuntimeError(f"gflow_b200: opacity must have {N} elements, got {opacity_c.numel()}")
        feature_c = _prep(feature, "feature", shape=(N, None))
        C = feature_c.shape[1]
        intr_c = _prep(intr, "intr", shape=(4,))
        extr_c = _prep(extr, "extr", shape=(3, 4))
        dev = _same_device(xyz_c, scale_c, rotate_c, opacity_c, feature_c, intr_c, extr_c)
        T = _grid(W, H)[0] * _grid(W, H)[1]
        key = (dev.index, N, W, H)
        lazy = _continues_a_loop(key, xyz_c) and bool(lazy) and _LAZY_K and key in _K_HINT
        with _on_device(dev):
            out = torch.empty(C, H, W, device=dev, dtype=torch.float32)
            kbuf, tbuf, aux, cap, K, ticket = _raster_forward(xyz_c, scale_c, rotate_c, opacity_c, feature_c, intr_c, extr_c,
                                                              N, C, W, H, bg, nearest, extent, dev, _capacity_for(key, N), out,
                                                              lazy)
            # backward buffers: d_cam 16 | d_rotate 4N | d_xyz 3N | d_scale 3N | d_opacity N | d_feature CN (the packed
            # per-Gaussian gradients live in a kept, self-cleaning block)
            bufs = (torch.empty(16, device=dev, dtype=torch.float32),
                    torch.empty((11 + C) * max(N, 1), device=dev, dtype=torch.float32))
        if lazy:  # validated when the backward starts; opacity / feature are kept for a corrective pass
            ctx.save_for_backward(xyz_c, scale_c, rotate_c, intr_c, extr_c, opacity_c, feature_c)
            ctx.pending = (ticket, key)
        else:
            ctx.save_for_backward(xyz_c, scale_c, rotate_c, intr_c, extr_c)
            ctx.pending = None
            _K_HINT[key] = K
        ctx.bufs = (kbuf, tbuf, aux) + bufs
        ctx.meta = (N, C, T, cap, W, H, bg, nearest, extent)
        return out

    @staticmethod
    def backward(ctx, g_out):
        xyz, scale, rotate, intr, extr = ctx.saved_tensors[:5]
        kbuf, tbuf, aux, d_cam, dbuf = ctx.bufs
        N, C, T, cap, W, H, bg, nearest, extent = ctx.meta
        dev = xyz.device
        with _on_device(dev):
            if ctx.pending is not None:  # lazy validation of the forward's speculative capacity
                ticket, key = ctx.pending
                ctx.pending = None
                k_host = _ctypes.c_int64(0)
                rc = _lib.gfb_wait_k_ticket(ticket, _ctypes.byref(k_host))
                if rc not in (0, capi.GFB_E_STALE):
                    capi.check(rc, "rasterization backward (K)")
                K = int(k_host.value) if rc == 0 else None
                if K is not None:
                    _K_HINT[key] = max(K, 1)
                if K is None or K > cap:  # corrective pass (or the ticket expired and nothing can be proven)
                    if K is not None:
                        _warnings.warn(_CLIPPED_WARNING, RuntimeWarning, stacklevel=2)
                    opacity_c, feature_c = ctx.saved_tensors[5:7]
                    scratch = torch.empty(C, H, W, device=dev, dtype=torch.float32)
                    cap2 = cap if K is None else K + K // 8 + 1024
                    kbuf, tbuf, aux, cap, K, _ = _raster_forward(xyz, scale, rotate, opacity_c, feature_c, intr, extr, N, C, W, H,
                                                                 bg, nearest, extent, dev, cap2, scratch, False)
                    _K_HINT[key] = max(K, 1)
                    ctx.bufs = (kbuf, tbuf, aux, d_cam, dbuf)
                    ctx.meta = (N, C, T, cap, W, H, bg, nearest, extent)
            if getattr(ctx, "bwd_done", False):  # retain_graph: earlier gradients alias the first buffers
                d_cam, dbuf = torch.empty_like(d_cam), torch.empty_like(dbuf)
            ctx.bwd_done = True
            g_out = _prep(g_out, "grad feature_map", shape=(C, H, W))
            dp = dbuf.data_ptr()
            kp, tp = kbuf.data_ptr(), tbuf.data_ptr()
            st = _stream()
            pkey, pack = _kept("grad_pack", dev, st, N, 48 * max(N, 1))
            with _KEEP_LOCK:
                rc = _lib.gfb_render_backward_keep(
                    xyz.data_ptr(), scale.data_ptr(), rotate.data_ptr(), intr.data_ptr(), extr.data_ptr(), N, W, H, C, bg,
                    nearest, extent, kp + 56 * cap, tp, cap, kp, kp + 32 * cap, aux.data_ptr(),
                    aux.data_ptr() + 4 * H * W, g_out.data_ptr(), pack.data_ptr(), d_cam.data_ptr(), dp + 16 * N, dp + 28 * N,
                    dp, dp + 40 * N, dp + 44 * N, st)
                if rc != 0:
                    _KEPT.pop(pkey, None)
            capi.check(rc, "rasterization backward")
        d_rotate = dbuf[:4 * N].view(N, 4)
        d_xyz = dbuf[4 * N:7 * N].view(N, 3)
        d_scale = dbuf[7 * N:10 * N].view(N, 3)
        d_opacity = dbuf[10 * N:11 * N].view(N, 1)
        d_feature = dbuf[11 * N:(11 + C) * N].view(N, C)
        return (d_xyz, d_scale, d_rotate, d_opacity, d_feature, d_cam[12:16], d_cam[:12].view(3, 4), None, None, None,
                None, None, None)


def rasterization_unfused(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg):
    """The five operators one after the other, exactly as /root/reference/gflow/utils/render.py:21-64 calls them."""
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles_touched = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    ids, tile_range = sort_gaussian(uv, depth, W, H, radius, tiles_touched)
    return alpha_blending(uv, conic, opacity, feature, ids, tile_range, bg, W, H)


def rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg):
    """Upstream's convenience `msplat.rasterization`: world-space Gaussians -> (C,H,W) image.

    Up to four channels run through the fused pipeline; more channels fall back to the operator chain.
    """
    if isinstance(feature, torch.Tensor) and feature.dim() == 2 and 1 <= feature.shape[1] <= 4:
        # a backward will follow (and validate K) only when autograd records this call
        lazy = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                               for t in (xyz, scale, rotate, opacity, feature, intr, extr))
        return _Rasterize.apply(xyz, scale, rotate, opacity, feature, intr, extr, int(W), int(H), float(bg), 0.2, 1.3, lazy)
    return rasterization_unfused(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg)


# --------------------------------------------------------------------------- C++ binding (optional)
# csrc/torch_ext.cpp implements the same operators as torch::autograd::Function in C++ on top of the
# same C ABI.  When the in-tree extension has been built (__graft_entry__.build()), the public names
# are rebound to it: identical kernels and results, ~5x less host time per operator.  The ctypes
# implementations above stay importable as `<name>_py` (tests run both).
def debug_set_k_hints(value: int) -> None:
    """Test hook: make every remembered K look like `value` (forces the GFB_E_CAPACITY retry)."""
    for k in list(_K_HINT):
        _K_HINT[k] = int(value)
    if _C is not None:
        _C.set_all_k_hints(int(value))


def _load_extension():
    import importlib.util
    import os

    from . import _build

    if os.environ.get("GFLOW_B200_NO_EXT") == "1" or not os.path.exists(_build.EXT_PATH) or _build.ext_needs_build():
        return None
    try:
        spec = importlib.util.spec_from_file_location(_build.EXT_NAME, _build.EXT_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception as e:  # pragma: no cover - a stale / ABI-mismatched build falls back to ctypes, loudly
        import warnings

        warnings.warn(f"gflow_b200: C++ binding present but not loadable ({e}); using the ctypes path")
        return None


project_point_py, compute_cov3d_py, ewa_project_py = project_point, compute_cov3d, ewa_project
sort_gaussian_py, alpha_blending_py, compute_sh_py = sort_gaussian, alpha_blending, compute_sh
rasterization_py, rasterization_unfused_py = rasterization, rasterization_unfused
_C = _load_extension()
BACKEND = "ctypes"
if _C is not None:
    BACKEND = "cpp_extension"

    def project_point(xyz, intr, extr, W, H, nearest=0.2, extent=1.3):  # noqa: F811
        """msplat.project_point -- /root/reference/gflow/utils/render.py:21-24, trainer.py:955."""
        return _C.project_point(xyz, intr, extr, int(W), int(H), float(nearest), float(extent))

    def compute_cov3d(scale, rotate, visible=None):  # noqa: F811
        """msplat.compute_cov3d -- /root/reference/gflow/utils/render.py:37-41."""
        return _C.compute_cov3d(scale, rotate, visible)

    def ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible=None):  # noqa: F811
        """msplat.ewa_project -- /root/reference/gflow/utils/render.py:44-49."""
        return _C.ewa_project(xyz, cov3d, intr, extr, uv, int(W), int(H), visible)

    def sort_gaussian(uv, depth, W, H, radius, tiles_touched):  # noqa: F811
        """msplat.sort_gaussian -- /root/reference/gflow/utils/render.py:52-54."""
        return _C.sort_gaussian(uv, depth, int(W), int(H), radius, tiles_touched)

    def alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, bg, W, H, ndc=None):  # noqa: F811
        """msplat.alpha_blending -- /root/reference/gflow/utils/render.py:58-64 (and 68-105, 148-154)."""
        return _C.alpha_blending(uv, conic, opacity, feature, gaussian_ids_sorted, tile_range, float(bg), int(W),
                                 int(H), ndc)

    def compute_sh(shs, dirs, visible=None):  # noqa: F811
        """msplat.compute_sh (north_star surface; not called by GFlow): shs (N,C,K), dirs (N,3)."""
        return _C.compute_sh(shs, dirs, visible)

    def rasterization_unfused(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg):  # noqa: F811
        """The five operators one after the other, exactly as /root/reference/gflow/utils/render.py:21-64 calls them."""
        uv, depth = project_point(xyz, intr, extr, W, H)
        visible = depth != 0
        cov3d = compute_cov3d(scale, rotate, visible)
        conic, radius, tiles_touched = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
        ids, tile_range = sort_gaussian(uv, depth, W, H, radius, tiles_touched)
        return alpha_blending(uv, conic, opacity, feature, ids, tile_range, bg, W, H)

    def rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg):  # noqa: F811
        """Upstream's convenience `msplat.rasterization`: fused pipeline up to four channels, operator chain beyond."""
        if isinstance(feature, torch.Tensor) and feature.dim() == 2 and 1 <= feature.shape[1] <= 4:
            return _C.rasterization_fused(xyz, scale, rotate, opacity, feature, intr, extr, int(W), int(H), float(bg))
        return rasterization_unfused(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg)
