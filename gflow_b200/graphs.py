"""The render step as ONE CUDA graph: forward + backward of `msplat.rasterization` captured once, replayed per step.

Eagerly, a render step costs ~125 us of host time (two trips through the autograd engine, nine launches) against
~105 us of kernels at BASELINE config 2: the GPU waits for the host.  `GraphedRenderStep` captures the six kernels of
gfb_render_forward / gfb_render_backward (C ABI, include/gflow_b200.h) into a CUDA graph over static buffers; a step is
then one cudaGraphLaunch and the GPU is the only limit.  The caller owns the parameters as the static tensors
`step.xyz ... step.extr` (update them in place, e.g. with an optimiser working on `step.parameters()`), fills
`step.g_image` with dL/d(image), calls `step()` and reads `step.image` / `step.grads`.

The intersection count K can drift while parameters move.  The graph is captured with room for `capacity`
intersections (default: K of a first eager pass + 25 %); the kernels clamp there, and `step.k()` / `step.check()` read
the K word of the last replay so the owner can re-capture when the scene outgrows the capacity.  There is no CPU
fallback: everything below needs the CUDA library.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import capi, ops


class GraphedRenderStep:
    def __init__(self, xyz, scale, rotate, opacity, feature, intr, extr, W: int, H: int, bg: float = 0.0,
                 capacity: Optional[int] = None, nearest: float = 0.2, extent: float = 1.3, backward: bool = True,
                 adopt_inputs: bool = False, capture: bool = True, grad_buffer: Optional[torch.Tensor] = None,
                 cam_buffer: Optional[torch.Tensor] = None):
        """adopt_inputs: use the given tensors themselves as the static inputs (they must be contiguous float32 CUDA
        tensors that stay alive) instead of cloning them.  capture=False: build the buffers only; the caller captures
        `enqueue()` inside a larger graph of its own.  grad_buffer: caller-provided flat float32 buffer of
        (11 + C) N elements that receives d_rotate | d_xyz | d_scale | d_opacity | d_feature; cam_buffer: likewise 16
        floats for d_extr (12) | d_intr (4)."""
        self.lib = capi.load()
        f32 = torch.float32
        own = (lambda t: t) if adopt_inputs else (lambda t: t.clone())
        self.xyz = own(ops._prep(xyz, "xyz", shape=(None, 3)).detach())
        N = self.N = self.xyz.shape[0]
        self.scale = own(ops._prep(scale, "scale", shape=(N, 3)).detach())
        self.rotate = own(ops._prep(rotate, "rotate", shape=(N, 4)).detach())
        self.opacity = own(ops._prep(opacity, "opacity").detach().reshape(-1))
        self.feature = own(ops._prep(feature, "feature", shape=(N, None)).detach())
        C = self.C = self.feature.shape[1]
        if not 1 <= C <= 4:
            raise RuntimeError("gflow_b200: the fused pipeline takes 1..4 feature channels")
        self.intr = own(ops._prep(intr, "intr", shape=(4,)).detach())
        self.extr = own(ops._prep(extr, "extr", shape=(3, 4)).detach())
        dev = self.dev = ops._same_device(self.xyz, self.scale, self.rotate, self.opacity, self.feature, self.intr, self.extr)
        self.W, self.H, self.bg, self.nearest, self.extent = int(W), int(H), float(bg), float(nearest), float(extent)
        self.with_backward = bool(backward)
        gx, gy = ops._grid(self.W, self.H)
        T = self.T = gx * gy
        with ops._on_device(dev):
            if capacity is None:  # one eager, synchronous pass tells K
                with torch.no_grad():
                    ops.rasterization_py(self.xyz, self.scale, self.rotate, self.opacity, self.feature, self.intr, self.extr,
                                         self.W, self.H, self.bg)
                k = ops._K_HINT[(dev.index, N, self.W, self.H)]
                capacity = k + k // 4 + 4096
            cap = self.capacity = int(capacity)
            self._gbuf = torch.empty(9 * max(N, 1), device=dev, dtype=f32)
            # tile_range (8T bytes) | control block: zero once, self-cleaning afterwards (gfb_render_forward_keep)
            self._tbuf = torch.zeros(8 * T + self.lib.gfb_render_control_bytes(self.W, self.H), device=dev, dtype=torch.uint8)
            self._kbuf = torch.empty(15 * max(cap, 1), device=dev, dtype=f32)
            self._aux = torch.empty(2, self.H, self.W, device=dev, dtype=f32)
            self.image = torch.empty(C, self.H, self.W, device=dev, dtype=f32)
            self.g_image = torch.zeros(C, self.H, self.W, device=dev, dtype=f32)
            self._grad_ws = torch.zeros(12 * N + 16, device=dev, dtype=f32)  # kept gradient pack (self-cleaning) | d_cam
            if cam_buffer is not None:
                if cam_buffer.numel() != 16 or cam_buffer.dtype != f32 or not cam_buffer.is_contiguous():
                    raise RuntimeError("gflow_b200: cam_buffer must be a contiguous float32 tensor of 16 elements")
                self._cam = cam_buffer
            else:
                self._cam = self._grad_ws[12 * N:12 * N + 16]
            if grad_buffer is not None:
                if grad_buffer.numel() != (11 + C) * N or grad_buffer.dtype != f32 or not grad_buffer.is_contiguous():
                    raise RuntimeError("gflow_b200: grad_buffer must be a contiguous float32 tensor of (11 + C) N elements")
                self._dbuf = grad_buffer
            else:
                self._dbuf = torch.empty((11 + C) * max(N, 1), device=dev, dtype=f32)
            d = self._dbuf
            self.grads: Dict[str, torch.Tensor] = {
                "rotate": d[:4 * N].view(N, 4), "xyz": d[4 * N:7 * N].view(N, 3), "scale": d[7 * N:10 * N].view(N, 3),
                "opacity": d[10 * N:11 * N].view(N, 1), "feature": d[11 * N:(11 + C) * N].view(N, C),
                "extr": self._cam[:12].view(3, 4), "intr": self._cam[12:16]}
            k_off = 8 * T + self.lib.gfb_render_control_k_offset(self.W, self.H)
            self._k_word = self._tbuf[k_off:k_off + 4].view(torch.int32)
            self.graph = self.graph_fwd = self.graph_bwd = None
            self.kernels_per_step = 0
            if capture:
                self.warm_up()
                self.graph = torch.cuda.CUDAGraph()  # forward + backward in one launch (g_image known beforehand)
                with torch.cuda.graph(self.graph):
                    self.enqueue()
                if self.with_backward:  # and as two graphs, for a loss computed from the image in between
                    self.graph_fwd, self.graph_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_fwd):
                        self.enqueue(backward=False)
                    with torch.cuda.graph(self.graph_bwd):
                        self.enqueue(forward=False)

    def warm_up(self, extra=None) -> None:
        """Two eager passes on a side stream: the first launches load modules, which cannot happen inside a capture."""
        dev = self.dev
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                n0 = self.lib.gfb_kernel_launch_count()
                self.enqueue()
                self.kernels_per_step = int(self.lib.gfb_kernel_launch_count() - n0)  # what one replay launches
                if extra is not None:
                    extra()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

    def enqueue(self, forward: bool = True, backward: bool = True) -> None:
        """Enqueues forward and / or backward on the current stream -- eagerly, or into the CUDA graph being captured."""
        N, C, W, H, cap, T = self.N, self.C, self.W, self.H, self.capacity, self.T
        gp, tp, kp = self._gbuf.data_ptr(), self._tbuf.data_ptr(), self._kbuf.data_ptr()
        st = ops._stream()
        if forward:
            self._enqueue_forward(gp, tp, kp, st)
        if not (backward and self.with_backward):
            return
        dp = self._dbuf.data_ptr()
        gw = self._grad_ws.data_ptr()
        capi.check(self.lib.gfb_render_backward_keep(
            self.xyz.data_ptr(), self.scale.data_ptr(), self.rotate.data_ptr(), self.intr.data_ptr(), self.extr.data_ptr(),
            N, W, H, C, self.bg, self.nearest, self.extent, kp + 56 * cap, tp, cap, kp, kp + 32 * cap, self._aux.data_ptr(),
            self._aux.data_ptr() + 4 * H * W, self.g_image.data_ptr(), gw, self._cam.data_ptr(), dp + 16 * N, dp + 28 * N, dp,
            dp + 40 * N, dp + 44 * N, st), "graphed rasterization backward")

    def _enqueue_forward(self, gp, tp, kp, st) -> None:
        N, C, W, H, cap, T = self.N, self.C, self.W, self.H, self.capacity, self.T
        capi.check(self.lib.gfb_render_forward_keep(
            self.xyz.data_ptr(), self.scale.data_ptr(), self.rotate.data_ptr(), self.opacity.data_ptr(),
            self.feature.data_ptr(), C, self.intr.data_ptr(), self.extr.data_ptr(), N, W, H, self.bg, self.nearest, self.extent,
            gp, gp + 16 * N, gp + 20 * N, gp + 32 * N, gp + 8 * N, tp + 8 * T, tp, cap, kp + 48 * cap, kp + 56 * cap,
            kp, kp + 32 * cap, self.image.data_ptr(), self._aux.data_ptr(), self._aux.data_ptr() + 4 * H * W, None, st),
            "graphed rasterization forward")

    def parameters(self):
        """The static input tensors (update them in place between replays)."""
        return [self.xyz, self.scale, self.rotate, self.opacity, self.feature, self.extr]

    def __call__(self) -> torch.Tensor:
        """One step: replays forward (+ backward against g_image).  Returns the static image tensor."""
        if self.graph is None:
            raise RuntimeError("gflow_b200: this GraphedRenderStep was built with capture=False; replay the owner's graph")
        self.graph.replay()
        return self.image

    def forward(self) -> torch.Tensor:
        """Forward only (its own graph): compute dL/d(image) from the returned image, write it into g_image, then backward()."""
        if self.graph_fwd is None:
            raise RuntimeError("gflow_b200: no separate forward graph (capture=False or backward=False)")
        self.graph_fwd.replay()
        return self.image

    def backward(self) -> Dict[str, torch.Tensor]:
        """Backward only, against the current contents of g_image and the last forward's buffers."""
        if self.graph_bwd is None:
            raise RuntimeError("gflow_b200: no separate backward graph (capture=False or backward=False)")
        self.graph_bwd.replay()
        return self.grads

    def k(self) -> int:
        """Intersection count of the last replay (synchronises)."""
        return int(self._k_word.item())

    def check(self) -> None:
        """Raises when the last replay needed more room than the graph was captured with (results were clamped)."""
        k = self.k()
        if k > self.capacity:
            raise RuntimeError(f"gflow_b200: the scene now has {k} tile intersections but the graph was captured for "
                               f"{self.capacity}; build a new GraphedRenderStep (capacity={k + k // 4 + 4096})")



class BatchedRenderStep:
    """A batch of F cameras over ONE set of Gaussians, forward + backward, the frames side by side.

    BASELINE config 5 names an "8-frame batch"; SURVEY.md 7 step 6 a batch-of-frames launch.  Every frame is a
    GraphedRenderStep over the SAME static parameter tensors (own camera, own buffers, own dL/d(image)), replayed on a
    stream of its own, then the per-frame gradients are summed by one kernel: the gradient of a loss summed over the
    views.  Frames of a batch do not depend on each other, so their kernels fill the SM time a lone frame leaves idle in
    every kernel's draining wave -- measured on a B200 at config 2: 14 300 frames/s for four frames side by side against
    9 400 for one frame after the other.  (One grid over the tiles of all frames would need the same per-frame
    intersection lists, sorts and buffers; streams get the overlap without a second set of kernels.)

    step = BatchedRenderStep(xyz, scale, rotate, opacity, rgb, intrs (F,4), extrs (F,3,4), W, H, bg)
    step.g_images[f].copy_(dL/d(image f));  step();  step.images[f], step.grads["xyz"], step.cam_grads[f]["extr"]
    """

    def __init__(self, xyz, scale, rotate, opacity, feature, intrs, extrs, W: int, H: int, bg: float = 0.0,
                 capacity: Optional[int] = None, nearest: float = 0.2, extent: float = 1.3):
        f32 = torch.float32
        self.xyz = ops._prep(xyz, "xyz", shape=(None, 3)).detach().clone()
        N = self.N = self.xyz.shape[0]
        self.scale = ops._prep(scale, "scale", shape=(N, 3)).detach().clone()
        self.rotate = ops._prep(rotate, "rotate", shape=(N, 4)).detach().clone()
        self.opacity = ops._prep(opacity, "opacity").detach().reshape(-1).clone()
        self.feature = ops._prep(feature, "feature", shape=(N, None)).detach().clone()
        C = self.C = self.feature.shape[1]
        self.intrs = ops._prep(intrs, "intrs", shape=(None, 4)).detach().clone()
        F = self.F = self.intrs.shape[0]
        self.extrs = ops._prep(extrs, "extrs", shape=(F, 3, 4)).detach().clone()
        if F < 1:
            raise RuntimeError("gflow_b200: BatchedRenderStep needs at least one camera")
        dev = self.dev = self.xyz.device
        self._all = torch.empty(F, (11 + C) * max(N, 1), device=dev, dtype=f32)  # per-frame gradients, one row each
        self._sum = torch.empty((11 + C) * max(N, 1), device=dev, dtype=f32)
        self.frames = []
        for f in range(F):
            st = GraphedRenderStep(self.xyz, self.scale, self.rotate, self.opacity, self.feature, self.intrs[f], self.extrs[f],
                                   W, H, bg, capacity=capacity, nearest=nearest, extent=extent, adopt_inputs=True,
                                   grad_buffer=self._all[f])
            self.frames.append(st)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
        self.images = [st.image for st in self.frames]
        self.g_images = [st.g_image for st in self.frames]
        self.cam_grads = [{"extr": st.grads["extr"], "intr": st.grads["intr"]} for st in self.frames]
        d = self._sum
        self.grads: Dict[str, torch.Tensor] = {
            "rotate": d[:4 * N].view(N, 4), "xyz": d[4 * N:7 * N].view(N, 3), "scale": d[7 * N:10 * N].view(N, 3),
            "opacity": d[10 * N:11 * N].view(N, 1), "feature": d[11 * N:(11 + C) * N].view(N, C)}
        torch.cuda.synchronize(dev)

    def parameters(self):
        """The shared static parameter tensors (update them in place between calls)."""
        return [self.xyz, self.scale, self.rotate, self.opacity, self.feature]

    def __call__(self, sum_gradients: bool = True):
        """Replays every frame (forward + backward against its g_image) side by side; returns the list of images."""
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            s.wait_stream(cur)
        for st, s in zip(self.frames, self.streams):
            with torch.cuda.stream(s):
                st.graph.replay()
        for s in self.streams:
            cur.wait_stream(s)
        if sum_gradients:
            torch.sum(self._all, dim=0, out=self._sum)
        return self.images

    def check(self) -> None:
        for st in self.frames:
            st.check()
