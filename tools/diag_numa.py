#!/usr/bin/env python
"""Does the placement of a rank's cores / pinned memory matter for host<->device traffic on a multi-GPU box?
torchrun --nproc-per-node N tools/diag_numa.py  (env MODE = identity | reversed | none)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mode = os.environ.get("MODE", "identity")
cores = sorted(os.sched_getaffinity(0))
per = len(cores) // world
if mode != "none":
    slot = lr if mode == "identity" else world - 1 - lr
    os.sched_setaffinity(0, set(cores[slot * per:(slot + 1) * per]))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 3360064 // 4
h_in = torch.empty(n).pin_memory(); h_in.fill_(1.0)          # first touch on the cores chosen above
h_out = torch.empty(n).pin_memory(); h_out.fill_(0.0)
d_in, d_out = torch.empty(n, device=dev), torch.ones(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(iters):
    for _ in range(iters):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    s1.synchronize(); s2.synchronize()
run(20)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(300); dt = time.perf_counter() - t0
gbs = 300 * 4 * n / dt / 1e9
t = torch.tensor([gbs], device=dev); g = [torch.zeros(1, device=dev) for _ in range(world)]
dist.all_gather(g, t)
if rank == 0:
    v = [round(float(x), 1) for x in g]
    print(f"mode {mode}: per-rank GB/s per direction {v}  sum {sum(v):.1f}", flush=True)
dist.destroy_process_group()
