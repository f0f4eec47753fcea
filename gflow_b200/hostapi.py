"""Render step with HOST buffers: the call a caller makes whose Gaussians live in host memory.

`HostRenderStep.submit(host_in, host_out)` takes the step's inputs from one pinned host block
(xyz | scale | rotate | opacity | feature | intr | extr, float32), runs `msplat.rasterization` forward + backward
against a fixed dL/d(image) and delivers the gradients + the loss into a pinned host block
(d_xyz | d_scale | d_rotate | d_opacity | d_feature | d_extr | loss).

The copies run on a stream of their own and the device buffers are double buffered, so the H2D copy of step i+1
and the D2H copy of step i-1 overlap the kernels of step i (round 1 ran copy -> compute -> copy serially on one
stream and lost 47 % against the device-resident step).  Nothing here is a CPU fallback: the work is done by the
CUDA kernels behind gflow_b200.ops.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import ops


class HostRenderStep:
    def __init__(self, N: int, W: int, H: int, feature_shape: Tuple[int, ...], g_image: torch.Tensor, bg: float = 0.0,
                 device=None, depth: int = 2, colour: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        """feature_shape: per-Gaussian shape of the colour input, (3,) for rgb or (3, 16) for degree-3 SH coefficients
        (then `colour(feature, xyz)` maps it to (N, C<=4) colours on the device).  g_image: dL/d(image) (C,H,W) on the
        device.  depth: number of in-flight steps (2 = double buffering)."""
        self.N, self.W, self.H, self.bg = int(N), int(W), int(H), float(bg)
        self.dev = torch.device(device) if device is not None else g_image.device
        if self.dev.type != "cuda":
            raise RuntimeError("gflow_b200: HostRenderStep needs a CUDA device (there is no CPU fallback)")
        self.g_image = g_image
        self.colour = colour
        fsz = 1
        for d in feature_shape:
            fsz *= int(d)
        self.in_shapes: List[Tuple[int, ...]] = [(N, 3), (N, 3), (N, 4), (N, 1), (N, *feature_shape), (4,), (3, 4)]
        self.out_shapes: List[Tuple[int, ...]] = [(N, 3), (N, 3), (N, 4), (N, 1), (N, *feature_shape), (3, 4), (1,)]
        self.in_sizes = [3 * N, 3 * N, 4 * N, N, fsz * N, 4, 12]
        self.out_sizes = [3 * N, 3 * N, 4 * N, N, fsz * N, 12, 1]
        self.in_offs = self._offsets(self.in_sizes)
        self.out_offs = self._offsets(self.out_sizes)
        self.h2d_bytes = 4 * self.in_offs[-1]
        self.d2h_bytes = 4 * self.out_offs[-1]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = []
        for _ in range(max(1, int(depth))):
            self.slots.append({
                "dev_in": torch.empty(self.in_offs[-1], dtype=torch.float32, device=self.dev),
                "dev_out": torch.empty(self.out_offs[-1], dtype=torch.float32, device=self.dev),
                "h2d": torch.cuda.Event(), "done": torch.cuda.Event(), "d2h": torch.cuda.Event(), "busy": False})
        self._next = 0

    @staticmethod
    def _offsets(sizes: Sequence[int]) -> List[int]:
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        return offs

    def host_input_block(self) -> torch.Tensor:
        return torch.empty(self.in_offs[-1], dtype=torch.float32).pin_memory()

    def host_output_block(self) -> torch.Tensor:
        return torch.empty(self.out_offs[-1], dtype=torch.float32).pin_memory()

    def pack_input(self, host_in: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
        for t, o, s in zip(tensors, self.in_offs, self.in_sizes):
            host_in[o:o + s].copy_(t.detach().reshape(-1).cpu())

    def unpack_output(self, host_out: torch.Tensor) -> List[torch.Tensor]:
        return [host_out[o:o + s].view(shape) for o, s, shape in zip(self.out_offs, self.out_sizes, self.out_shapes)]

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor) -> None:
        """Enqueue one step.  host_in / host_out must be pinned; host_out is valid after wait()."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise RuntimeError("gflow_b200: HostRenderStep.submit needs pinned host blocks (host_input_block / host_output_block)")
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        compute = torch.cuda.current_stream(self.dev)
        if slot["busy"]:
            # the slot's device buffers are free once its previous D2H has been issued after its compute; ordering on
            # the copy stream guarantees that, and the compute stream must not overwrite dev_out before that D2H ran
            compute.wait_event(slot["d2h"])
        with torch.cuda.stream(self.copy_stream):
            slot["dev_in"].copy_(host_in, non_blocking=True)
            slot["h2d"].record(self.copy_stream)
        compute.wait_event(slot["h2d"])
        dv = [slot["dev_in"][self.in_offs[i]:self.in_offs[i + 1]].view(self.in_shapes[i]) for i in range(7)]
        ps = [d.detach().requires_grad_(True) for d in dv[:5]]
        ex = dv[6].detach().requires_grad_(True)
        col = self.colour(ps[4], ps[0]) if self.colour is not None else ps[4]
        img = ops.rasterization(ps[0], ps[1], ps[2], ps[3], col, dv[5], ex, self.W, self.H, self.bg)
        loss = (img.detach() * self.g_image).sum()  # loss = sum(out * G): dL/d(out) = G goes to autograd directly
        img.backward(self.g_image)
        out = slot["dev_out"]
        for i, t in enumerate([p.grad for p in ps] + [ex.grad, loss]):
            out[self.out_offs[i]:self.out_offs[i + 1]].copy_(t.reshape(-1))
        slot["done"].record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["done"])
            host_out.copy_(out, non_blocking=True)
            slot["d2h"].record(self.copy_stream)
        slot["busy"] = True

    def wait(self) -> None:
        """Blocks until every submitted step's results have landed in their host blocks."""
        for slot in self.slots:
            if slot["busy"]:
                slot["d2h"].synchronize()
