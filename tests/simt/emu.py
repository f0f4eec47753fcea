"""TEST-ONLY driver of the emulated kernel library (tests/simt/build_emu.py): calls the C ABI of
include/gflow_b200.h with HOST tensors, mirroring what gflow_b200/ops.py does with device tensors."""
from __future__ import annotations

import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import build_emu  # noqa: E402

_lib = None


def load():
    global _lib
    if _lib is None:
        from gflow_b200.capi import SIGNATURES

        # The emulated suite pins the fused pipeline's ids / tile_range to the oracle's 3-sigma enumeration bit for
        # bit, so it runs with tile culling off unless a test asks for it (tests/simt/tight_tiles_check.py runs the
        # shipped default, GFB_TIGHT_TILES=1, in a process of its own; the switch is read once per process).
        os.environ.setdefault("GFB_TIGHT_TILES", "0")
        # GFB_EMU_LIB: an alternative build of the same sources (the UBSan one, tests/test_simt_kernels.py)
        path = os.environ.get("GFB_EMU_LIB") or build_emu.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def set_schedule(mode: int, seed: int = 0):
    """Fiber sweep order inside a CTA: 0 ascending, 1 descending, 2 pseudo-random per sweep.  A correct
    kernel gives the same results under all of them; a missing wait / barrier usually does not."""
    lib = load()
    lib.gfb_emu_set_schedule.argtypes = [ctypes.c_int, ctypes.c_uint]
    lib.gfb_emu_set_schedule.restype = None
    lib.gfb_emu_set_schedule(mode, seed)


def ok(rc, what):
    assert rc == 0, f"{what}: rc {rc}"


def p(t):
    return None if t is None else t.data_ptr()


def f32(*shape):
    return torch.full(shape, float("nan"), dtype=torch.float32)


def i32(*shape):
    return torch.full(shape, -12345, dtype=torch.int32)


def operator_chain(sc, Gimg, C=3, feature=None):
    """project_point .. alpha_blending one by one + every backward, the way ops.py drives them."""
    L = load()
    N, W, H = sc.xyz.shape[0], sc.W, sc.H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    feature = sc.rgb if feature is None else feature
    C = feature.shape[1]
    xyz, scale, rot, op = sc.xyz.contiguous(), sc.scale.contiguous(), sc.rotate.contiguous(), sc.opacity.contiguous()
    intr, extr = sc.intr.contiguous(), sc.extr.contiguous()
    uv, depth = f32(N, 2), f32(N, 1)
    ok(L.gfb_project_point_fwd(p(xyz), p(intr), p(extr), N, W, H, 0.2, 1.3, p(uv), p(depth), None), "project")
    vis = (depth != 0).to(torch.uint8).contiguous()
    cov = f32(N, 6)
    ok(L.gfb_compute_cov3d_fwd(p(scale), p(rot), p(vis), N, p(cov), None), "cov3d")
    conic, radius, tiles = f32(N, 3), i32(N, 1), i32(N, 1)
    ok(L.gfb_ewa_project_fwd(p(xyz), p(cov), p(intr), p(extr), p(uv), N, W, H, p(vis), p(conic), p(radius), p(tiles),
                             None), "ewa")
    cap = max(1, int(tiles.sum()))
    tile_ws = torch.zeros(L.gfb_sort_tile_workspace_bytes(W, H), dtype=torch.uint8)
    keys = torch.zeros(L.gfb_sort_workspace_bytes(cap), dtype=torch.uint8)
    ids, rng = i32(cap), i32(T, 2)
    K = ctypes.c_int64(-1)
    ok(L.gfb_sort_gaussian(p(uv), p(depth), p(radius), p(tiles), N, W, H, p(tile_ws), cap, p(keys), p(ids), p(rng),
                           ctypes.addressof(K), None), "sort")
    K = int(K.value)
    # the kept-workspace variant: zero block in, same result, zero block out
    tile_ws.zero_()
    ids_k, rng_k, K_k = i32(cap), i32(T, 2), ctypes.c_int64(-1)
    ok(L.gfb_sort_gaussian_keep(p(uv), p(depth), p(radius), p(tiles), N, W, H, p(tile_ws), cap, p(keys), p(ids_k),
                                p(rng_k), ctypes.addressof(K_k), None), "sort (kept workspace)")
    assert int(K_k.value) == K and torch.equal(ids_k[:K], ids[:K]) and torch.equal(rng_k, rng)
    n_clean = (tile_ws.numel() // 4 - 2) // 2 + 1  # counters[T R] | ticket; the offsets behind them are plain outputs
    assert not tile_ws.view(torch.int32)[:n_clean].any(), "gfb_sort_gaussian_keep left its counters dirty"
    ids = ids[:K].contiguous()
    geom = torch.zeros(max(K, 1) * 8, dtype=torch.float32)
    out, fT, nc = f32(C, H, W), f32(H, W), i32(H, W)
    ok(L.gfb_blend_pack_geometry(p(uv), p(conic), p(op), p(ids), K, p(geom), None), "pack geometry")
    d_uv, d_conic, d_op, d_feat = f32(N, 2), f32(N, 3), f32(N, 1), f32(N, C)
    gp = torch.zeros(N * 12, dtype=torch.float32)  # kept over the channel groups: GFB_UNPACK_CLEAR (2) hands it back zeroed
    for gi, c0 in enumerate(range(0, C, 4)):
        cg = min(4, C - c0)
        feat = torch.zeros(max(K, 1) * 4, dtype=torch.float32)
        ok(L.gfb_blend_pack_feature(p(feature.contiguous()), C, c0, cg, p(ids), K, p(feat), None), "pack feature")
        if gi == 0:  # the one-launch variant writes the same two streams
            geom2, feat2 = torch.zeros_like(geom), torch.zeros_like(feat)
            ok(L.gfb_blend_pack_geometry_feature(p(uv), p(conic), p(op), p(feature.contiguous()), C, 0, cg, p(ids), K,
                                                 p(geom2), p(feat2), None), "pack geometry + feature")
            assert torch.equal(geom2.view(torch.int32), geom.view(torch.int32)) and torch.equal(feat2, feat)
        ok(L.gfb_alpha_blending_fwd(p(geom), p(feat), K, p(rng), C, c0, cg, sc.bg, W, H, p(out), p(fT), p(nc), None),
           "blend fwd")
        ok(L.gfb_alpha_blending_bwd(p(geom), p(feat), K, p(ids), p(rng), C, c0, cg, sc.bg, W, H, p(fT), p(nc),
                                    p(Gimg.contiguous()), p(gp), None), "blend bwd")
        ok(L.gfb_blend_unpack_grads(p(gp), N, C, c0, cg, p(d_uv), p(d_conic), p(d_op), p(d_feat), 2 | int(gi > 0), None),
           "unpack")
        assert not gp.any(), "gfb_blend_unpack_grads(GFB_UNPACK_CLEAR) left the gradient pack dirty"
    d_xyz_e, d_cov, d_cam_e = f32(N, 3), f32(N, 6), f32(16)
    ok(L.gfb_ewa_project_bwd(p(xyz), p(cov), p(intr), p(extr), p(uv), N, W, H, p(vis), p(d_conic), p(d_xyz_e), p(d_cov),
                             p(d_cam_e), None), "ewa bwd")
    d_scale, d_rot = f32(N, 3), f32(N, 4)
    ok(L.gfb_compute_cov3d_bwd(p(scale), p(rot), p(vis), N, p(d_cov), p(d_scale), p(d_rot), None), "cov3d bwd")
    d_xyz_p, d_cam_p = f32(N, 3), f32(16)
    ok(L.gfb_project_point_bwd(p(xyz), p(intr), p(extr), N, W, H, 0.2, 1.3, p(d_uv), None, p(d_xyz_p), p(d_cam_p), None),
       "project bwd")
    d_cam = d_cam_e + d_cam_p
    return dict(uv=uv, depth=depth, cov3d=cov, conic=conic, radius=radius, tiles=tiles, ids=ids, tile_range=rng, K=K,
                image=out, final_T=fT, n_contrib=nc,
                grads=dict(xyz=d_xyz_e + d_xyz_p, scale=d_scale, rotate=d_rot, opacity=d_op, feature=d_feat,
                           extr=d_cam[:12].reshape(3, 4), intr=d_cam[12:]))


def fused_pipeline(sc, Gimg, feature=None, capacity=None, kept=None):
    """gfb_render_forward / gfb_render_backward (msplat.rasterization).  kept: dict holding the caller-kept,
    self-cleaning workspaces across calls ("control", "pack"): the *_keep entry points are used then."""
    L = load()
    N, W, H = sc.xyz.shape[0], sc.W, sc.H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    feature = (sc.rgb if feature is None else feature).contiguous()
    C = feature.shape[1]
    xyz, scale, rot, op = sc.xyz.contiguous(), sc.scale.contiguous(), sc.rotate.contiguous(), sc.opacity.contiguous()
    intr, extr = sc.intr.contiguous(), sc.extr.contiguous()
    uv, depth, conic, radius = f32(N, 2), f32(N, 1), f32(N, 3), i32(N, 1)
    rect = torch.zeros(max(N, 1) * 8, dtype=torch.uint8)
    if kept is not None:
        ctrl = kept.setdefault("control", torch.zeros(L.gfb_render_control_bytes(W, H), dtype=torch.uint8))
    else:
        ctrl = torch.full((L.gfb_render_control_bytes(W, H),), 0x5A, dtype=torch.uint8)
    rng = i32(T, 2)
    cap = int(capacity) if capacity is not None else 64 * max(N, 1)
    keys = torch.zeros(max(cap, 1) * 8, dtype=torch.uint8)
    ids = i32(max(cap, 1))
    geom = torch.zeros(max(cap, 1) * 8, dtype=torch.float32)
    feat = torch.zeros(max(cap, 1) * 4, dtype=torch.float32)
    out, fT, nc = f32(C, H, W), f32(H, W), i32(H, W)
    K = ctypes.c_int64(-1)
    fwd = L.gfb_render_forward_keep if kept is not None else L.gfb_render_forward
    rc = fwd(p(xyz), p(scale), p(rot), p(op), p(feature), C, p(intr), p(extr), N, W, H, sc.bg, 0.2, 1.3,
                              p(uv), p(depth), p(conic), p(radius), p(rect), p(ctrl), p(rng), cap, p(keys), p(ids),
                              p(geom), p(feat), p(out), p(fT), p(nc), ctypes.addressof(K), None)
    K = int(K.value)
    if rc != 0:
        return dict(rc=rc, K=K)
    d_xyz, d_scale, d_rot, d_op, d_feat = f32(N, 3), f32(N, 3), f32(N, 4), f32(N, 1), f32(N, C)
    if kept is not None:
        pack = kept.setdefault("pack", torch.zeros(max(N, 1) * 12, dtype=torch.float32))
        d_cam = torch.full((16,), float("nan"), dtype=torch.float32)
        ok(L.gfb_render_backward_keep(p(xyz), p(scale), p(rot), p(intr), p(extr), N, W, H, C, sc.bg, 0.2, 1.3, p(ids), p(rng),
                                      cap, p(geom), p(feat), p(fT), p(nc), p(Gimg.contiguous()), p(pack), p(d_cam), p(d_xyz),
                                      p(d_scale), p(d_rot), p(d_op), p(d_feat), None), "render backward (kept workspaces)")
    else:
        gws = torch.full((L.gfb_render_grad_bytes(N) // 4,), float("nan"), dtype=torch.float32)
        ok(L.gfb_render_backward(p(xyz), p(scale), p(rot), p(intr), p(extr), N, W, H, C, sc.bg, 0.2, 1.3, p(ids), p(rng), cap,
                                 p(geom), p(feat), p(fT), p(nc), p(Gimg.contiguous()), p(gws), p(d_xyz), p(d_scale), p(d_rot),
                                 p(d_op), p(d_feat), None), "render backward")
        d_cam = gws[N * 12:N * 12 + 16]
    return dict(rc=0, uv=uv, depth=depth, conic=conic, radius=radius, ids=ids[:K].clone(), tile_range=rng, K=K, image=out,
                final_T=fT, n_contrib=nc,
                grads=dict(xyz=d_xyz, scale=d_scale, rotate=d_rot, opacity=d_op, feature=d_feat,
                           extr=d_cam[:12].reshape(3, 4).clone(), intr=d_cam[12:].clone()))



def fit_loop_class():
    """gflow_b200.fit.NativeFitLoop (the product's host driver of csrc/fit.cu) re-pointed at the emulated
    kernel library and host tensors, so its chunking / overflow / workspace logic runs in the CPU suite."""
    import contextlib

    from gflow_b200.fit import NativeFitLoop

    class EmulatedFitLoop(NativeFitLoop):
        def _require_device(self, dev):
            assert dev.type == "cpu", "the emulated loop runs on host tensors"

        def _library(self):
            return load()

        def _stream(self):
            return 0

        def _device_guard(self):
            return contextlib.nullcontext()

        def _make_densifier(self):
            return densifier_class()(self.W, self.H, self.dev)

    return EmulatedFitLoop


def densifier_class():
    """gflow_b200.densify.Densifier re-pointed at the emulated library and host tensors."""
    from gflow_b200.densify import Densifier

    class EmulatedDensifier(Densifier):
        def _require_device(self, dev):
            assert dev.type == "cpu"

        def _library(self):
            return load()

        def _stream(self):
            return 0

    return EmulatedDensifier
