"""Frame-state files in GFlow's checkpoint format, and their packed wire form for frame-sharded runs.

GFlow writes one `ckpt/<frame>.tar` per optimised frame (/root/reference/gflow/trainer.py:252-272) and reads
them back in its benchmark / viewer (/root/reference/gflow/trainer.py:274-288,
/root/reference/gflow/viewer.py, /root/reference/gflow/benchmark.py):

    torch.save({"attributes": {xyz, scale, rotate, opacity, rgb: raw nn.Parameters},
                "intr": (4,), "extr": (3,4) world->camera, "still_mask": (N,) bool | None,
                "move_seg": (H,W) uint8 ndarray | None, "last_uv": (N,2) | None,
                "width": W, "height": H[, "pose_list": [...]]}, path)

`save_frame` / `load_frame` keep exactly that dictionary (same keys, raw pre-activation attributes, extr as
rendered by pose -> [R|t]) so files move freely between this package and the reference's tools.  Loading
re-derives the 7-float pose from extr the way `load_camera(extr=...)` does (trainer.py:170-179: rotation
matrix -> xyzw unit quaternion; `signed_log1p` is the identity, utils/__init__.py:11-15).

`to_wire` / `from_wire` flatten a frame state into one float32 tensor ( header | N x 14 attributes | extras )
-- what the ranks of a frame-sharded run exchange in a single NCCL broadcast / gather (gflow_b200.frames).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import frames as _frames

ATTRS = tuple(k for k, _ in _frames.STATE_KEYS)
_WIRE_MAGIC = 0x47464C57  # "GFLW"
_WIRE_HEADER = 8           # magic, version, N, W, H, has_still, has_last_uv, reserved


@dataclass
class FrameState:
    attributes: Dict[str, torch.Tensor]        # raw (pre-activation) attributes, (N,3) (N,3) (N,4) (N,1) (N,3)
    intr: torch.Tensor                          # (4,) fx fy cx cy
    extr: torch.Tensor                          # (3,4) world -> camera
    width: int
    height: int
    still_mask: Optional[torch.Tensor] = None   # (N,) bool
    move_seg: Optional[np.ndarray] = None       # (H,W) uint8
    last_uv: Optional[torch.Tensor] = None      # (N,2)
    pose_list: Optional[list] = None

    @property
    def num_points(self) -> int:
        return int(self.attributes["xyz"].shape[0])

    def pose(self) -> torch.Tensor:
        """(qx, qy, qz, qw, tx, ty, tz) as load_camera(extr=...) derives it (trainer.py:170-179)."""
        from .fit import extr_to_pose

        return extr_to_pose(self.extr.detach().float().cpu()).to(self.extr.device)


def _check_attributes(attrs: Dict[str, torch.Tensor]) -> int:
    missing = [k for k in ATTRS if k not in attrs]
    if missing:
        raise ValueError(f"checkpoint attributes lack {missing}")
    N = int(attrs["xyz"].shape[0])
    for k, w in _frames.STATE_KEYS:
        if tuple(attrs[k].shape) != (N, w):
            raise ValueError(f"attribute {k!r} must have shape ({N}, {w}), got {tuple(attrs[k].shape)}")
    return N


def save_frame(path: str, state: FrameState) -> str:
    """Writes `state` in the reference's checkpoint layout (trainer.py:252-272)."""
    _check_attributes(state.attributes)
    ckpt = {
        "attributes": {k: torch.nn.Parameter(state.attributes[k].detach().clone(), requires_grad=True) for k in ATTRS},
        "intr": state.intr.detach().clone(),
        "extr": state.extr.detach().clone(),
        "still_mask": None if state.still_mask is None else state.still_mask.detach().clone(),
        "move_seg": state.move_seg,
        "last_uv": None if state.last_uv is None else state.last_uv.detach().clone(),
        "width": int(state.width),
        "height": int(state.height),
    }
    if state.pose_list is not None:
        ckpt["pose_list"] = state.pose_list
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    torch.save(ckpt, path)
    return path


def load_frame(path: str, device=None) -> FrameState:
    """Reads a checkpoint written by GFlow or by save_frame.  The file is a pickle (the reference's format
    holds nn.Parameters and a numpy array), so only load files you trust."""
    ckpt = torch.load(path, map_location=device, weights_only=False)
    for key in ("attributes", "intr", "extr"):
        if key not in ckpt:
            raise ValueError(f"{path}: not a GFlow checkpoint (no {key!r})")
    attrs = {k: v.detach() for k, v in ckpt["attributes"].items()}
    _check_attributes(attrs)
    extr = torch.as_tensor(ckpt["extr"]).detach()
    if tuple(extr.shape) == (4, 4):
        extr = extr[:3]
    if tuple(extr.shape) != (3, 4):
        raise ValueError(f"{path}: extr must be (3,4), got {tuple(extr.shape)}")
    return FrameState(attributes=attrs, intr=torch.as_tensor(ckpt["intr"]).detach(), extr=extr,
                      width=int(ckpt.get("width", 0)), height=int(ckpt.get("height", 0)),
                      still_mask=ckpt.get("still_mask"), move_seg=ckpt.get("move_seg"), last_uv=ckpt.get("last_uv"),
                      pose_list=ckpt.get("pose_list"))


def to_wire(state: FrameState) -> torch.Tensor:
    """One float32 tensor: header(8) | intr(4) | extr(12) | N x 14 attributes | still_mask(N)? | last_uv(2N)?.
    move_seg (a per-frame output image) and pose_list do not travel."""
    N = _check_attributes(state.attributes)
    dev = state.attributes["xyz"].device
    header = torch.tensor([_WIRE_MAGIC, 1, N, state.width, state.height, int(state.still_mask is not None),
                           int(state.last_uv is not None), 0], dtype=torch.int32, device=dev).view(torch.float32)
    parts = [header, state.intr.reshape(-1).float().to(dev), state.extr.reshape(-1).float().to(dev),
             _frames.pack_state(state.attributes).reshape(-1)]
    if state.still_mask is not None:
        if state.still_mask.numel() != N:
            raise ValueError("still_mask must have one entry per Gaussian")
        parts.append(state.still_mask.to(dev).float().reshape(-1))
    if state.last_uv is not None:
        if tuple(state.last_uv.shape) != (N, 2):
            raise ValueError("last_uv must be (N,2)")
        parts.append(state.last_uv.to(dev).float().reshape(-1))
    return torch.cat(parts)


def wire_numel(N: int, has_still: bool, has_last_uv: bool) -> int:
    return _WIRE_HEADER + 16 + _frames.STATE_WIDTH * N + (N if has_still else 0) + (2 * N if has_last_uv else 0)


def from_wire(buf: torch.Tensor) -> FrameState:
    if buf.dtype != torch.float32 or buf.dim() != 1 or buf.numel() < _WIRE_HEADER + 16:
        raise ValueError("wire buffer must be a flat float32 tensor")
    h = buf[:_WIRE_HEADER].view(torch.int32).tolist()
    if h[0] != _WIRE_MAGIC or h[1] != 1:
        raise ValueError("not a gflow_b200 frame-state buffer (bad magic / version)")
    N, W, H, has_still, has_uv = h[2], h[3], h[4], bool(h[5]), bool(h[6])
    if buf.numel() != wire_numel(N, has_still, has_uv):
        raise ValueError(f"wire buffer has {buf.numel()} floats, expected {wire_numel(N, has_still, has_uv)}")
    o = _WIRE_HEADER
    intr, extr = buf[o:o + 4].clone(), buf[o + 4:o + 16].reshape(3, 4).clone()
    o += 16
    attrs = _frames.unpack_state(buf[o:o + _frames.STATE_WIDTH * N].reshape(N, _frames.STATE_WIDTH))
    o += _frames.STATE_WIDTH * N
    still = last_uv = None
    if has_still:
        still = buf[o:o + N] != 0
        o += N
    if has_uv:
        last_uv = buf[o:o + 2 * N].reshape(N, 2).clone()
    return FrameState(attributes=attrs, intr=intr, extr=extr, width=W, height=H, still_mask=still, last_uv=last_uv)
