#!/usr/bin/env python
"""Where the first frames.broadcast_state / gather_frames of a job spend their time (per rank, host clock with a device
sync after every stage):  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/diag_bcast.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from gflow_b200 import frames

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
frames.warm_up(dev) if hasattr(frames, "warm_up") else dist.barrier()
N = 60000


def stage(name, fn, log):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    log.append((name, 1e3 * (time.perf_counter() - t0)))
    return r


for rep in range(3):
    log = []
    state = {k: torch.rand(N, w, device=dev) for k, w in frames.STATE_KEYS} if rank == 0 else None
    stage("barrier", dist.barrier, log)
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    flat = stage("pack", lambda: frames.pack_state(state).to(dev), log) if rank == 0 else None
    if rank == 0:
        n[0] = flat.shape[0]
    stage("bcast_n", lambda: dist.broadcast(n, src=0), log)
    cnt = stage("item", lambda: int(n.item()), log)
    if rank != 0:
        flat = stage("alloc", lambda: torch.empty(cnt, frames.STATE_WIDTH, device=dev), log)
    stage("bcast_flat", lambda: dist.broadcast(flat, src=0), log)
    stage("unpack", lambda: frames.unpack_state(flat), log)
    stage("barrier2", dist.barrier, log)
    whole = stage("broadcast_state()", lambda: frames.broadcast_state(state, src=0, device=dev), log)
    img, pose = torch.rand(3, 480, 854, device=dev), torch.rand(3, 4, device=dev)
    stage("barrier3", dist.barrier, log)
    stage("gather_frames()", lambda: frames.gather_frames(img, pose, dst=0), log)
    if rank in (0, world - 1):
        print(f"rank {rank} rep {rep}: " + "  ".join(f"{k} {v:.2f}" for k, v in log), flush=True)
dist.destroy_process_group()
