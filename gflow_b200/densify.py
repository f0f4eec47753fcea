"""Error-driven densification on the device (csrc/densify.cu) -- the host side of
SimpleGaussian.densify_by_pixels (/root/reference/gflow/trainer.py:878-939) without its GPU -> CPU -> GPU
round trip.  One 4-byte read (the number of mask pixels, which sizes the new tensors) is the only
synchronisation."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import capi, ops


class Densifier:
    """Workspace + calls for one image size.  CUDA tensors only (no CPU fallback)."""

    def __init__(self, W: int, H: int, device):
        self.W, self.H, self.dev = int(W), int(H), torch.device(device)
        self._require_device(self.dev)
        self.lib = self._library()
        self.ws = torch.empty(self.lib.gfb_densify_workspace_bytes(self.W, self.H), dtype=torch.uint8, device=self.dev)
        self.error_map = torch.empty(self.H, self.W, dtype=torch.float32, device=self.dev)

    # hooks (the CPU test-suite re-points them at the emulated kernel library)
    def _require_device(self, dev) -> None:
        if dev.type != "cuda":
            raise RuntimeError("gflow_b200: densification needs CUDA tensors (no CPU fallback exists)")

    def _library(self):
        return capi.load()

    def _stream(self) -> int:
        return ops._stream()

    def rgb_error_map(self, rendered: torch.Tensor, gt_image: torch.Tensor, pixel_mask: Optional[torch.Tensor] = None):
        """loss_rgb_pixel (H,W) of trainer.py:457 from a (>=3,H,W) render and the (H,W,3) target."""
        pm = None if pixel_mask is None else pixel_mask.to(torch.uint8).contiguous()
        capi.check(self.lib.gfb_rgb_error_map(rendered.data_ptr(), gt_image.data_ptr(), ops._ptr(pm), self.W, self.H,
                                              self.error_map.data_ptr(), self._stream()), "rgb error map")
        return self.error_map

    def sample(self, error_map: torch.Tensor, gt_image: torch.Tensor, gt_depth: torch.Tensor, intr: torch.Tensor,
               extr: torch.Tensor, num_points: int, error_threshold: float = 1e-3, percent: float = 0.1,
               mask: Optional[torch.Tensor] = None, seed: int = 0) -> Optional[Dict[str, torch.Tensor]]:
        """Returns the new raw attributes {xyz, scale, rotate, opacity, rgb, pixels} or None when
        int(num_points * mask_ratio * percent) == 0 (trainer.py:900-902)."""
        f32 = dict(dtype=torch.float32, device=self.dev)
        err = error_map.detach().to(**f32).reshape(self.H, self.W).contiguous()
        m8 = None if mask is None else mask.detach().to(self.dev).reshape(self.H, self.W).to(torch.uint8).contiguous()
        st = self._stream()
        capi.check(self.lib.gfb_densify_prepare(err.data_ptr(), ops._ptr(m8), self.W, self.H, float(error_threshold),
                                                self.ws.data_ptr(), st), "densify prepare")
        mask_count = int(self.ws[:32].view(torch.int32)[1])  # the one synchronising read
        count = int(int(num_points) * (mask_count / float(self.W * self.H)) * float(percent))
        if count <= 0:
            return None
        new = {k: torch.empty(count, w, **f32) for k, w in (("xyz", 3), ("scale", 3), ("rotate", 4), ("opacity", 1), ("rgb", 3))}
        pixels = torch.empty(count, dtype=torch.int32, device=self.dev)
        gi, gd = gt_image.detach().to(**f32).contiguous(), gt_depth.detach().to(**f32).contiguous()
        it, ex = intr.detach().to(**f32).contiguous(), extr.detach().to(**f32).contiguous()
        capi.check(self.lib.gfb_densify_sample(self.ws.data_ptr(), gi.data_ptr(), gd.data_ptr(), it.data_ptr(), ex.data_ptr(),
                                               self.W, self.H, count, int(num_points), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                               new["xyz"].data_ptr(), new["scale"].data_ptr(), new["rotate"].data_ptr(),
                                               new["opacity"].data_ptr(), new["rgb"].data_ptr(), pixels.data_ptr(), st),
                   "densify sample")
        new["pixels"] = pixels
        return new
