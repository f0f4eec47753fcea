"""Frame-sharded execution of the per-frame optimisation loop.

The reference fits frames sequentially in one process (/root/reference/gflow/fit_video.py:242-349);
per-frame problems are independent given a start state, so this module shards *frames* across
ranks (one process per GPU, torch.distributed over NCCL/NVLink):

  * one broadcast of the packed Gaussian state (N first, because N is dynamic) at sequence start,
  * no collective inside the iteration loop,
  * one gather of per-frame outputs (rendered frame, pose) at the end.

Both collectives are a few MB -- microseconds on NVSwitch -- so plain NCCL calls are the right
tool; there is no compute step to fuse them with.  The packing / sharding logic is backend
agnostic and is covered by world_size-2 gloo tests on CPU (tests/test_frames_gloo.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

STATE_KEYS: Tuple[Tuple[str, int], ...] = (("xyz", 3), ("scale", 3), ("rotate", 4), ("opacity", 1), ("rgb", 3))
STATE_WIDTH = sum(w for _, w in STATE_KEYS)  # 14 floats per Gaussian


def shard_frames(num_frames: int, world: int, rank: int) -> range:
    """Contiguous chunk of frame indices owned by `rank` (ceil(F/R) per rank, last ranks may be short)."""
    per = (num_frames + world - 1) // world
    lo = min(num_frames, rank * per)
    hi = min(num_frames, lo + per)
    return range(lo, hi)


def warm_up(device, payload_floats: int = 1 << 20) -> None:
    """Bring the communicator up before the first payload collective: NCCL connects channels lazily, per collective
    and per message-size class, and the first frames.* call of a job otherwise pays for that (hundreds of ms at 8
    ranks).  One barrier, one small and one payload-sized broadcast / all_gather, then the call patterns below once."""
    dist.barrier()
    small = torch.ones(1, device=device)
    dist.all_reduce(small)
    dist.broadcast(small, src=0)
    big = torch.zeros(int(payload_floats), device=device)
    dist.broadcast(big, src=0)
    dist.all_gather([torch.empty_like(big) for _ in range(dist.get_world_size())], big)
    tiny = {k: torch.zeros(8, w, device=device) for k, w in STATE_KEYS}
    broadcast_state(tiny if dist.get_rank() == 0 else None, src=0, device=device)
    gather_frames(torch.zeros(3, 8, 8, device=device), torch.zeros(3, 4, device=device), dst=0)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def pack_state(state: Dict[str, torch.Tensor]) -> torch.Tensor:
    """(N,14) float32: xyz | scale | rotate | opacity | rgb (checkpoint attribute order of
    /root/reference/gflow/trainer.py:81-88)."""
    N = state["xyz"].shape[0]
    for k, w in STATE_KEYS:
        if tuple(state[k].shape) != (N, w):
            raise ValueError(f"state[{k!r}] must have shape ({N}, {w}), got {tuple(state[k].shape)}")
    return torch.cat([state[k].detach().to(torch.float32) for k, _ in STATE_KEYS], dim=1).contiguous()


def unpack_state(flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, c = {}, 0
    for k, w in STATE_KEYS:
        out[k] = flat[:, c:c + w].contiguous()
        c += w
    return out


def broadcast_state(state: Optional[Dict[str, torch.Tensor]], src: int = 0, device=None) -> Dict[str, torch.Tensor]:
    """Broadcast the Gaussian state from `src`; other ranks pass None.  Two messages: N, then N x 14 floats."""
    rank = dist.get_rank()
    if device is None:
        device = state["xyz"].device if state is not None else torch.device("cpu")
    n = torch.zeros(1, dtype=torch.int64, device=device)
    flat = None
    if rank == src:
        flat = pack_state(state).to(device)
        n[0] = flat.shape[0]
    dist.broadcast(n, src=src)
    if rank != src:
        flat = torch.empty(int(n.item()), STATE_WIDTH, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src)
    return unpack_state(flat)


def gather_frames(image: torch.Tensor, pose: torch.Tensor, dst: int = 0) -> Optional[List[Tuple[torch.Tensor, torch.Tensor]]]:
    """Gather one (C,H,W) frame and its (3,4) pose per rank on `dst` (one all_gather of a packed buffer)."""
    world = dist.get_world_size()
    payload = torch.cat([image.detach().reshape(-1).to(torch.float32),
                         pose.detach().reshape(-1).to(torch.float32)]).contiguous()
    bufs = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(bufs, payload)
    if dist.get_rank() != dst:
        return None
    n_img = image.numel()
    return [(b[:n_img].reshape(image.shape), b[n_img:].reshape(pose.shape)) for b in bufs]


def broadcast_frame_state(state, src: int = 0, device=None):
    """Broadcast a whole checkpoint.FrameState (attributes + camera + still mask + last_uv) from `src` as ONE
    packed float32 buffer (checkpoint.to_wire); other ranks pass None.  Two messages: length, then payload."""
    from . import checkpoint as _ckpt

    rank = dist.get_rank()
    if device is None:
        device = state.attributes["xyz"].device if state is not None else torch.device("cpu")
    n = torch.zeros(1, dtype=torch.int64, device=device)
    buf = None
    if rank == src:
        buf = _ckpt.to_wire(state).to(device)
        n[0] = buf.numel()
    dist.broadcast(n, src=src)
    if rank != src:
        buf = torch.empty(int(n.item()), dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    return _ckpt.from_wire(buf)
