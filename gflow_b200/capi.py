"""ctypes binding of libgflow_b200.so (the C ABI of include/gflow_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library cannot be built / loaded the
import raises, and every op raises when handed a tensor that is not on a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

from . import _build

_lib = None
GFB_E_CAPACITY, GFB_E_STALE, GFB_E_NOTREADY = -3, -4, -5

# symbol -> (restype, argtypes); mirrors include/gflow_b200.h declaration by declaration
P, I, F, L = c_void_p, c_int, c_float, c_int64
SIGNATURES = {
    "gfb_version": (c_int, []),
    "gfb_build_arch": (ctypes.c_char_p, []),
    "gfb_error_string": (ctypes.c_char_p, [I]),
    "gfb_kernel_launch_count": (c_int64, []),
    "gfb_project_point_fwd": (I, [P, P, P, I, I, I, F, F, P, P, P]),
    "gfb_project_point_bwd": (I, [P, P, P, I, I, I, F, F, P, P, P, P, P]),
    "gfb_compute_cov3d_fwd": (I, [P, P, P, I, P, P]),
    "gfb_compute_cov3d_bwd": (I, [P, P, P, I, P, P, P, P]),
    "gfb_ewa_project_fwd": (I, [P, P, P, P, P, I, I, I, P, P, P, P, P]),
    "gfb_ewa_project_bwd": (I, [P, P, P, P, P, I, I, I, P, P, P, P, P, P]),
    "gfb_compute_sh_fwd": (I, [P, P, P, I, I, I, P, P]),
    "gfb_compute_sh_bwd": (I, [P, P, P, I, I, I, P, P, P, P]),
    "gfb_sort_workspace_bytes": (c_size_t, [L]),
    "gfb_sort_tile_workspace_bytes": (c_size_t, [I, I]),
    "gfb_sort_gaussian": (I, [P, P, P, P, I, I, I, P, L, P, P, P, P, P]),
    "gfb_sort_gaussian_keep": (I, [P, P, P, P, I, I, I, P, L, P, P, P, P, P]),
    "gfb_render_control_bytes": (c_size_t, [I, I]),
    "gfb_render_control_k_offset": (c_size_t, [I, I]),
    "gfb_wait_k": (I, [P]),
    "gfb_k_ticket": (c_int64, []),
    "gfb_wait_k_ticket": (I, [L, P]),
    "gfb_query_k_ticket": (I, [L, P]),
    "gfb_render_forward": (I, [P, P, P, P, P, I, P, P, I, I, I, F, F, F, P, P, P, P, P, P, P, L, P, P, P, P, P, P, P,
                               P, P]),
    "gfb_render_forward_keep": (I, [P, P, P, P, P, I, P, P, I, I, I, F, F, F, P, P, P, P, P, P, P, L, P, P, P, P, P, P, P,
                                    P, P]),
    "gfb_render_grad_bytes": (c_size_t, [I]),
    "gfb_render_backward": (I, [P, P, P, P, P, I, I, I, I, F, F, F, P, P, L, P, P, P, P, P, P, P, P, P, P, P, P]),
    "gfb_render_backward_keep": (I, [P, P, P, P, P, I, I, I, I, F, F, F, P, P, L, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    "gfb_blend_geometry_stream_bytes": (c_size_t, [L]),
    "gfb_blend_feature_stream_bytes": (c_size_t, [L]),
    "gfb_blend_grad_pack_bytes": (c_size_t, [I]),
    "gfb_blend_pack_geometry": (I, [P, P, P, P, L, P, P]),
    "gfb_blend_pack_feature": (I, [P, I, I, I, P, L, P, P]),
    "gfb_blend_pack_geometry_feature": (I, [P, P, P, P, I, I, I, P, L, P, P, P]),
    "gfb_alpha_blending_fwd": (I, [P, P, L, P, I, I, I, F, I, I, P, P, P, P]),
    "gfb_alpha_blending_bwd": (I, [P, P, L, P, P, I, I, I, F, I, I, P, P, P, P, P]),
    "gfb_blend_unpack_grads": (I, [P, I, I, I, I, P, P, P, P, I, P]),
    "gfb_fit_get_layout": (I, [I, I, I, L, I, P]),
    "gfb_fit_sub_workspace_bytes": (c_size_t, [I, I, I, L]),
    "gfb_fit_init": (I, [P, P, L, I, P]),
    "gfb_fit_iterate": (I, [P, P, L, I, I, I, P]),
    "gfb_rgb_error_map": (I, [P, P, P, I, I, P, P]),
    "gfb_densify_workspace_bytes": (c_size_t, [I, I]),
    "gfb_densify_prepare": (I, [P, P, I, I, F, P, P]),
    "gfb_densify_sample": (I, [P, P, P, P, P, I, I, I, I, ctypes.c_uint64, P, P, P, P, P, P, P]),
    "gfb_hostpipe_create": (I, [I, P]),
    "gfb_hostpipe_submit": (I, [P, I, P, P, c_size_t, P, P, P, P, c_size_t]),
    "gfb_hostpipe_wait": (I, [P]),
    "gfb_hostpipe_destroy": (I, [P]),
}


class FitProblem(ctypes.Structure):
    """struct gfb_fit_problem of include/gflow_b200.h, field by field."""
    _fields_ = [(n, P) for n in ("xyz", "scale", "rotate", "opacity", "rgb", "pose", "depth_ab", "intr", "gt_image",
                                 "gt_depth", "pixel_mask", "still_mask", "scale_sel", "still_ref", "still_sel", "flow_target",
                                 "flow_sel", "sub_xyz", "sub_scale", "sub_rotate", "sub_opacity", "sub_rgb",
                                 "dyn_mask", "sub_workspace", "dbg_grads", "dbg_act")] + \
               [(n, ctypes.c_int32) for n in ("N", "W", "H", "n_still", "n_still_ref", "still_count", "n_flow",
                                              "flow_count", "sub_N", "sub_capacity", "total_iters", "camera_only", "freeze_rgb", "use_ssim", "adam_t0",
                                              "constant_lr", "freeze_camera")] + \
               [(n, F) for n in ("bg", "nearest", "extent", "lr", "lr_camera", "lambda_rgb", "lambda_depth",
                                 "lambda_var", "lambda_scale", "lambda_still", "lambda_flow", "beta1", "beta2", "eps",
                                 "depth_den_min")]


class FitLayout(ctypes.Structure):
    """struct gfb_fit_layout of include/gflow_b200.h."""
    _fields_ = [(n, c_size_t) for n in ("status", "loss_hist", "cam", "adam_m", "adam_v", "uv", "depth", "conic",
                                        "radius", "tile_range", "ids", "out", "g_out", "total")]


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing or stale).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("GFLOW_B200_NO_BUILD") != "1" and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB_PATH):
        raise ImportError(f"gflow_b200: CUDA library missing at {_build.LIB_PATH}; run __graft_entry__.build()")
    # GFLOW_B200_LIB: another build of the same C ABI (A/B of whole-library variants, tools/build_variants.py); use it
    # together with GFLOW_B200_NO_EXT=1, the C++ binding is linked against the default library
    lib = ctypes.CDLL(os.environ.get("GFLOW_B200_LIB") or _build.LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gfb_error_string(rc).decode()
        raise RuntimeError(f"gflow_b200 {what} failed: {msg} (code {rc})")
