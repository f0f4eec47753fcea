#!/usr/bin/env python
"""BASELINE config 3 / 4: the 300-iteration per-frame Adam loop, single GPU or frame-sharded.

  python tools/bench_fit.py [--iters 300] [--frames 1] [--fused | --native] [--ssim]
  torchrun --nproc-per-node N tools/bench_fit.py --frames 48      (config 4, frame-sharded)

Prints one JSON line: frames x iterations per second (max wall time over ranks, including the
broadcast of the state and the gather of the per-frame outputs).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gflow_b200 import fit, frames  # noqa: E402
from gflow_b200.synthetic import make_camera, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--points", type=int, default=60000)
ap.add_argument("--fused", action="store_true")
ap.add_argument("--no-depth", action="store_true")
ap.add_argument("--size", type=int, nargs=2, default=[854, 480], metavar=("W", "H"))
ap.add_argument("--native", action="store_true", help="whole iteration in csrc/fit.cu (no autograd / torch.optim)")
ap.add_argument("--concurrent", type=int, default=1, help="native: frames of one GPU run side by side, one stream each")
ap.add_argument("--ssim", action="store_true", help="loss_rgb = mse + (1 - SSIM) as in gflow/trainer.py:459-462")
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
W, H = args.size
sc = make_scene(args.points, W, H, seed=0, profile="gflow")
raw = {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
       "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
       "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}
state0 = {k: v.to(dev) for k, v in raw.items()} if rank == 0 else None
if world == 1:
    state0 = {k: v.to(dev) for k, v in raw.items()}


class Targets:
    """Per-frame synthetic priors: the scene rendered from a slightly different camera, plus its depth."""

    def __len__(self):
        return args.frames

    def __call__(self, i):
        gen = torch.Generator().manual_seed(100 + i)
        _, extr = make_camera(W, H, gen)
        pose = fit.extr_to_pose(extr)
        f = fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), pose.to(dev), W, H)
        with torch.no_grad():
            img, dmap, _ = f.render(0.0, want_depth=True)
        return img.permute(1, 2, 0).contiguous(), (None if args.no_depth else dmap.permute(1, 2, 0).contiguous()), \
            fit.extr_to_pose(sc.extr)


cfg = fit.FitConfig(iterations=args.iters, lr=4e-3, lr_camera=1e-3, lambda_depth=0.0 if args.no_depth else 0.1,
                    fused=args.fused, native=args.native, use_ssim=args.ssim)
targets = Targets()
# warm-up (allocator, K hints, NCCL)
fit.FrameFitter({k: v.to(dev) for k, v in raw.items()}, sc.intr.to(dev), fit.extr_to_pose(sc.extr).to(dev), W, H).train(
    *[t for t in targets(0)[:2]], fit.FitConfig(iterations=5, lambda_depth=cfg.lambda_depth, fused=args.fused, native=args.native,
                                        use_ssim=args.ssim))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
results, gathered = fit.fit_sequence_sharded(state0, sc.intr, targets, W, H, cfg, dev, concurrent_frames=args.concurrent)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
if rank == 0:
    first = results[min(results)]
    print(json.dumps({"metric": "per-frame Adam loop, frames x iterations / s", "value": args.frames * args.iters / dt,
                      "unit": "iters/s", "n_gpus": world, "frames": args.frames, "iterations": args.iters,
                      "points": args.points, "resolution": [W, H], "seconds": dt, "fused": args.fused, "native": args.native, "ssim": args.ssim, "concurrent_frames": args.concurrent,
                      "depth_loss": not args.no_depth, "loss_first": first.losses[0], "loss_last": first.losses[-1],
                      "frames_per_rank": [len(frames.shard_frames(args.frames, world, r)) for r in range(world)]}))
if world > 1:
    dist.destroy_process_group()
