"""Render step with HOST buffers: the call a caller makes whose Gaussians live in host memory.

`HostRenderStep.submit(host_in, host_out)` takes the step's inputs from one pinned host block
(layout below), runs `msplat.rasterization` forward + backward against a fixed dL/d(image) and delivers the
gradients + the loss into a pinned host block.

The copies run on two streams of their own (one per direction: PCIe is full duplex and the GPU has a copy engine for
each) and the device buffers are double buffered, so the H2D copy of step i+1 and the D2H copy of step i-1 overlap
the kernels of step i (round 1 ran copy -> compute -> copy serially on one
stream and lost 47 % against the device-resident step).  With plain per-Gaussian colours the compute part of a slot
is ONE CUDA graph (gflow_b200.graphs.GraphedRenderStep over the slot's buffers + the loss) and the whole step -- both
copies, the graph launch and the events between them -- is ONE call into the library (gfb_hostpipe_submit,
csrc/hostpipe.cu): issued from Python, those ten calls cost the host more than the kernels take.  With a `colour`
callback (e.g. spherical harmonics through autograd) the step runs eagerly on torch streams.
Nothing here is a CPU fallback: the work is done by the CUDA kernels behind gflow_b200.ops.

Host block layouts (float32):
  input   xyz 3N | scale 3N | rotate 4N | opacity N | feature F N | intr 4 | extr 12
  output  d_rotate 4N | d_xyz 3N | d_scale 3N | d_opacity N | d_feature F N | d_extr 12 | d_intr 4 | loss 1
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import capi, ops


class HostRenderStep:
    def __init__(self, N: int, W: int, H: int, feature_shape: Tuple[int, ...], g_image: torch.Tensor, bg: float = 0.0,
                 device=None, depth: int = 2, colour: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None,
                 capacity: Optional[int] = None, sample_input: Optional[torch.Tensor] = None, concurrent: bool = True):
        """feature_shape: per-Gaussian shape of the colour input, (3,) for rgb or (3, 16) for degree-3 SH coefficients
        (then `colour(feature, xyz)` maps it to (N, C<=4) colours on the device).  g_image: dL/d(image) (C,H,W) on the
        device.  depth: number of in-flight steps (2 = double buffering).  sample_input: a representative host input
        block; needed by the graph path to size the intersection capacity (or pass `capacity`).  concurrent (graph path):
        every slot computes on a stream of its own, so neighbouring steps overlap; g_image must then not be modified
        while steps are in flight (wait() first)."""
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
        self.in_sizes = [3 * N, 3 * N, 4 * N, N, fsz * N, 4, 12]
        self.out_names = ["rotate", "xyz", "scale", "opacity", "feature", "extr", "intr", "loss"]
        self.out_shapes: List[Tuple[int, ...]] = [(N, 4), (N, 3), (N, 3), (N, 1), (N, *feature_shape), (3, 4), (4,), (1,)]
        self.out_sizes = [4 * N, 3 * N, 3 * N, N, fsz * N, 12, 4, 1]
        self.in_offs = self._offsets(self.in_sizes)
        self.out_offs = self._offsets(self.out_sizes)
        self.h2d_bytes = 4 * self.in_offs[-1]
        self.d2h_bytes = 4 * self.out_offs[-1]
        self.h2d_stream = torch.cuda.Stream(device=self.dev)
        self.d2h_stream = torch.cuda.Stream(device=self.dev)
        self.graphed = colour is None and len(feature_shape) == 1 and 1 <= int(feature_shape[0]) <= 4
        self.slots = []
        for _ in range(max(1, int(depth))):
            slot = {"dev_in": torch.empty(self.in_offs[-1], dtype=torch.float32, device=self.dev),
                    "dev_out": torch.zeros(self.out_offs[-1], dtype=torch.float32, device=self.dev),
                    "h2d": torch.cuda.Event(), "done": torch.cuda.Event(), "d2h": torch.cuda.Event(), "busy": False}
            if self.graphed:
                if sample_input is not None:
                    slot["dev_in"].copy_(sample_input)
                elif capacity is None:
                    raise RuntimeError("gflow_b200: the graph path of HostRenderStep needs `sample_input` or `capacity`")
                self._capture(slot, capacity)
                capacity = slot["step"].capacity  # the other slots reuse the first one's capacity
            self.slots.append(slot)
        self._next = 0
        self._pipe = None
        if self.graphed:
            self._lib = capi.load()
            pipe = ctypes.c_void_p()
            with torch.cuda.device(self.dev):
                capi.check(self._lib.gfb_hostpipe_create(len(self.slots), ctypes.byref(pipe)), "host pipe")
            self._pipe = pipe
            for slot in self.slots:
                slot["exec"] = int(slot["graph"].raw_cuda_graph_exec())
                # own compute stream per slot: the kernels of step i+1 fill the SMs the last wave of step i leaves idle
                slot["stream"] = torch.cuda.Stream(device=self.dev) if concurrent else None
            torch.cuda.synchronize(self.dev)

    def __del__(self):
        pipe, self._pipe = getattr(self, "_pipe", None), None
        if pipe is not None:
            try:
                self._lib.gfb_hostpipe_wait(pipe)
                self._lib.gfb_hostpipe_destroy(pipe)
            except Exception:  # interpreter shutdown
                pass

    def _capture(self, slot, capacity) -> None:
        """The slot's compute as one CUDA graph: render step over the slot's input views, all gradients (per-Gaussian
        and camera) straight into the slot's output block, the loss appended by one captured dot product."""
        from .graphs import GraphedRenderStep

        N = self.N
        dv = [slot["dev_in"][self.in_offs[i]:self.in_offs[i + 1]].view(self.in_shapes[i]) for i in range(7)]
        out = slot["dev_out"]
        n_grad = self.out_offs[5]
        step = GraphedRenderStep(dv[0], dv[1], dv[2], dv[3], dv[4], dv[5], dv[6], self.W, self.H, self.bg, capacity=capacity,
                                 adopt_inputs=True, capture=False, grad_buffer=out[:n_grad],
                                 cam_buffer=out[n_grad:n_grad + 16])  # the kernels write straight into the output block
        step.g_image = self.g_image  # dL/d(image) is shared by the slots (read only)
        loss = out[n_grad + 16]      # 0-dim view of the block's last word

        def tail():  # loss = sum(image * G) as one dot product, no temporary
            torch.dot(step.image.view(-1), self.g_image.reshape(-1), out=loss)

        step.warm_up(tail)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step.enqueue()
            tail()
        slot["step"], slot["graph"] = step, graph

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

    def unpack_output(self, host_out: torch.Tensor) -> dict:
        """{"rotate", "xyz", "scale", "opacity", "feature", "extr", "intr", "loss"} views of a host output block."""
        return {n: host_out[o:o + s].view(shape) for n, o, s, shape in zip(self.out_names, self.out_offs, self.out_sizes, self.out_shapes)}

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor) -> None:
        """Enqueue one step.  host_in / host_out must be pinned; host_out is valid after wait()."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise RuntimeError("gflow_b200: HostRenderStep.submit needs pinned host blocks (host_input_block / host_output_block)")
        index = self._next
        slot = self.slots[index]
        self._next = (self._next + 1) % len(self.slots)
        compute = torch.cuda.current_stream(self.dev)
        if self._pipe is not None:  # graph path: the whole step is one library call
            capi.check(self._lib.gfb_hostpipe_submit(self._pipe, index, slot["dev_in"].data_ptr(), host_in.data_ptr(),
                                                     self.h2d_bytes, slot["exec"],
                                                     (slot["stream"] or compute).cuda_stream, host_out.data_ptr(),
                                                     slot["dev_out"].data_ptr(), self.d2h_bytes), "host pipe submit")
            slot["busy"] = True
            return
        with torch.cuda.stream(self.h2d_stream):
            if slot["busy"]:
                self.h2d_stream.wait_event(slot["done"])  # dev_in is still read by the slot's previous compute
            slot["dev_in"].copy_(host_in, non_blocking=True)
            slot["h2d"].record(self.h2d_stream)
        compute.wait_event(slot["h2d"])
        if slot["busy"]:
            compute.wait_event(slot["d2h"])  # do not overwrite dev_out before its previous D2H has run
        if self.graphed:
            slot["graph"].replay()
        else:
            self._eager(slot)
        slot["done"].record(compute)
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(slot["done"])
            host_out.copy_(slot["dev_out"], non_blocking=True)
            slot["d2h"].record(self.d2h_stream)
        slot["busy"] = True

    def _eager(self, slot) -> None:
        dv = [slot["dev_in"][self.in_offs[i]:self.in_offs[i + 1]].view(self.in_shapes[i]) for i in range(7)]
        ps = [d.detach().requires_grad_(True) for d in dv[:5]]
        ex = dv[6].detach().requires_grad_(True)
        col = self.colour(ps[4], ps[0]) if self.colour is not None else ps[4]
        img = ops.rasterization(ps[0], ps[1], ps[2], ps[3], col, dv[5], ex, self.W, self.H, self.bg)
        loss = (img.detach() * self.g_image).sum()  # loss = sum(out * G): dL/d(out) = G goes to autograd directly
        img.backward(self.g_image)
        out = slot["dev_out"]
        grads = {"rotate": ps[2].grad, "xyz": ps[0].grad, "scale": ps[1].grad, "opacity": ps[3].grad, "feature": ps[4].grad,
                 "extr": ex.grad, "loss": loss}
        for i, n in enumerate(self.out_names):
            if n in grads and grads[n] is not None:
                out[self.out_offs[i]:self.out_offs[i + 1]].copy_(grads[n].reshape(-1))

    def check(self) -> None:
        """Graph path: raises when a slot's last step needed more intersections than it was captured for."""
        if self.graphed:
            for slot in self.slots:
                slot["step"].check()

    def wait(self) -> None:
        """Blocks until every submitted step's results have landed in their host blocks."""
        if self._pipe is not None:
            capi.check(self._lib.gfb_hostpipe_wait(self._pipe), "host pipe wait")
            return
        for slot in self.slots:
            if slot["busy"]:
                slot["d2h"].synchronize()
