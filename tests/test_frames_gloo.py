"""world_size-2 gloo tests (CPU) of the frame-sharding plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gflow_b200 import frames


def test_shard_frames_partition():
    for F in (1, 7, 48, 50, 56):
        for R in (1, 2, 4, 8):
            chunks = [list(frames.shard_frames(F, R, r)) for r in range(R)]
            flat = [i for c in chunks for i in c]
            assert flat == list(range(F)), (F, R)
            assert max(len(c) for c in chunks) == (F + R - 1) // R
    assert list(frames.shard_frames(50, 8, 7)) == [49]  # 7/7/.../1: ideal speed-up 50/7


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    st = {k: torch.randn(37, w, generator=g) for k, w in frames.STATE_KEYS}
    flat = frames.pack_state(st)
    assert flat.shape == (37, 14)
    back = frames.unpack_state(flat)
    assert all(torch.equal(back[k], st[k]) for k, _ in frames.STATE_KEYS)
    with pytest.raises(ValueError):
        frames.pack_state({**st, "rgb": torch.zeros(36, 3)})


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        ref = {k: torch.randn(101, w, generator=g) for k, w in frames.STATE_KEYS}
        got = frames.broadcast_state(ref if rank == 0 else None, src=0, device=torch.device("cpu"))
        ok = all(torch.equal(got[k], ref[k]) for k, _ in frames.STATE_KEYS)
        img = torch.full((3, 4, 5), float(rank))
        pose = torch.arange(12.0).reshape(3, 4) + rank
        out = frames.gather_frames(img, pose, dst=0)
        if rank == 0:
            ok = ok and len(out) == world and all(
                torch.equal(out[r][0], torch.full((3, 4, 5), float(r))) and
                torch.equal(out[r][1], torch.arange(12.0).reshape(3, 4) + r) for r in range(world))
        else:
            ok = ok and out is None
        owned = list(frames.shard_frames(5, world, rank))
        q.put((rank, ok, owned))
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] and res[1][1], res
    assert res[0][2] == [0, 1, 2] and res[1][2] == [3, 4]
