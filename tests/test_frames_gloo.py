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
        frames.warm_up(torch.device("cpu"), payload_floats=64)  # the bring-up sequence bench.py runs before its clock
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


# ----------------------------------------------------------------------------- sharded fit_video loop (world 2, gloo)
def _seq_worker(rank, world, port, q):
    """Each rank fits its chunk of a 4-frame synthetic video with SequenceFitter; the native iteration runs through the
    SIMT shim (no GPU here), the collectives through gloo."""
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.join(here, "simt"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        import fit_check
        from gflow_b200 import fit, sequence

        fit.NativeFitLoop = emu.fit_loop_class()
        fit.FrameFitter.render = lambda self, bg=0.0, want_depth=True, with_depth=False: (torch.full((3, self.H, self.W), float(rank)), None, None)
        W, H, N = 48, 32, 150
        sc, raw, pose, gt_image, gt_depth = fit_check.make_problem(N, W, H, seed=50)
        move = torch.zeros(H, W, dtype=torch.bool)
        move[8:20, 10:30] = True
        flow = torch.zeros(H, W, 2)
        flow[..., 0] = 1.0

        def frame_inputs(i):
            return dict(image=torch.roll(gt_image, shifts=i, dims=1), depth=gt_depth, move_mask=move, flow=flow,
                        occ_mask=torch.zeros(H, W, 1))

        cfg = sequence.SequenceConfig(num_points=N, iterations_first=2, iterations_camera=1, iterations_after=2, densify_interval=0,
                                      densify_times=0, densify_interval_after=0, densify_times_after=0, lambda_var=0.1, native=True)
        outs, gathered = sequence.fit_video_sharded(raw if rank == 0 else None, sc.intr, pose, 4, frame_inputs, W, H, cfg,
                                                    torch.device("cpu"))
        ok = sorted(outs) == ([0, 1] if rank == 0 else [2, 3])
        # frame 0: first-frame recipe; frame 2 heads rank 1's chunk without a camera prior: camera-only stage, then the
        # first-frame recipe, flagged as a chunk head; the others: camera-only + full stages
        want = {0: {"first"}, 2: {"first", "camera"}, 1: {"camera", "all"}, 3: {"camera", "all"}}
        ok = ok and all(set(o.losses) == want[i] for i, o in outs.items())
        ok = ok and all(o.chunk_head == (i == 2) for i, o in outs.items())
        ok = ok and all(len(v) > 0 and all(x == x for x in v) for o in outs.values() for v in o.losses.values())
        if rank == 0:
            ok = ok and gathered is not None and len(gathered) == world
            ok = ok and all(float(gathered[r][0].mean()) == float(r) and gathered[r][1].shape == (3, 4) for r in range(world))
        else:
            ok = ok and gathered is None
        q.put((rank, bool(ok)))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def test_sharded_fit_video_loop_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seq_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
