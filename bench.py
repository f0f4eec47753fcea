#!/usr/bin/env python
"""bench.py -- splat forward+backward iterations/s (BASELINE.json metric) on 1..8 B200.

A "step" is one render step of SURVEY.md 8d unit (ii): project_point + compute_cov3d +
ewa_project + sort_gaussian + alpha_blending(C=3) and the whole backward chain, called
through the msplat operator surface (gflow_b200.ops -> C ABI -> sm_100a kernels) on the
synthetic 60 000-Gaussian / 854x480 scene (BASELINE config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

What one line reports:
  value            device-resident render step as ONE CUDA graph (gflow_b200.GraphedRenderStep: the five kernels of
                   gfb_render_forward_keep + gfb_render_backward_keep, one cudaGraphLaunch per step); BLOCKS blocks of K
                   steps, each block bracketed by a barrier + synchronize, CUDA events around every step (L2 flushed in
                   between), max over ranks per block, MEDIAN block reported.  With SH colours (cfg5): the eager step
  eager            the same step as the eager autograd call msplat.rasterization(...).backward(): same kernels, but
                   host-bound (107 us of host work per 106 us of kernels), so it follows the box's CPU
  graphed          = value's measurement, plus frames_in_flight: four independent graphed steps on four streams
  e2e              the same step through gflow_b200.hostapi.HostRenderStep with pinned HOST buffers: H2D of the
                   inputs and D2H of gradients + loss inside the timed region, --e2e-depth steps in flight
  reference_gpu    probe for a real MSplat on the box (baseline/_ref, site-packages); timed in the same harness if found
  operator_chain   the step through the five operators one by one (render.py:21-64 pattern)
  gflow_iteration  SURVEY 8d unit (iii): the render_multiple call pattern (four blends over one sort) + the rgb and
                   depth backward (render.py:6-108, trainer.py:404-533)
  roofline         alpha-blending backward timed alone, against the HBM roofline and the issue roofline
  sequence         BASELINE config 4: 48 synthetic frames, 300-iteration native Adam loop each, sharded by frame
                   across the ranks, wall clock including the one broadcast and the one gather
  cpu_baseline     the CPU port of the path (oracle/splat_oracle.c, OpenMP on all host cores), N = 1 only

N > 1 is launched by torchrun (one rank per GPU); every rank works on its own frames (frame sharding):
one NCCL broadcast of the Gaussian state before and one gather of per-frame outputs after, no collective inside.

--impl reference times the CPU port on rank 0 with all host cores: msplat ships no CPU kernels and is not installable
here, so the oracle port is the reference arm (kind "port").  At N > 1 it prints the same single-host figure.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]:
    # the CPU arm uses every host core it may run on; torchrun exports OMP_NUM_THREADS=1, and the OpenMP runtime
    # reads the variable when it is loaded -- so this has to happen before anything is imported
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))

import argparse  # noqa: E402
import ctypes  # noqa: E402
import importlib.util  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import time  # noqa: E402

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "splat fwd+bwd iters/sec @60k Gaussians/480p"
UNIT = "iters/s"
WORKLOADS = {"cfg1": (1_000, 256, 256), "cfg2": (60_000, 854, 480), "cfg5": (200_000, 1280, 720)}
BLOCKS = 10
SEQ_FRAMES, SEQ_ITERS = 48, 300
FRAMES_IN_FLIGHT = 4  # graphed.frames_in_flight: independent frames of one GPU side by side
SEQ_RUNS = 5
SEQ_CONCURRENT = 3  # independent frames a rank fits side by side (one stream each): 48 / N frames per rank divide evenly


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--profile", default="synthetic", choices=["synthetic", "gflow"])
    ap.add_argument("--blocks", type=int, default=BLOCKS, help="timed blocks of --steps steps; the median block is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-fit-loop", action="store_true", help="skip the config-3 Adam-loop section")
    ap.add_argument("--no-sequence", action="store_true", help="skip the config-4 frame-sharded sequence section")
    ap.add_argument("--no-proxy", action="store_true", help="skip the eager-PyTorch-on-CUDA proxy baseline")
    ap.add_argument("--quick", action="store_true", help="value / e2e / roofline only")
    ap.add_argument("--e2e-depth", type=int, default=6, help="steps in flight in the e2e leg (HostRenderStep depth)")
    return ap.parse_args()


def synthetic_module():
    """gflow_b200/synthetic.py loaded as a plain file: the reference arm must not import the package, whose __init__
    loads the CUDA library."""
    name = "gfb_synthetic_standalone"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "gflow_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst: kernel timed alone)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(N, K, P, T, C=3):
    """SURVEY.md 8d / BASELINE.md per-op compulsory traffic."""
    key_bytes = (32 + max(1, (T - 1).bit_length()) + 7) // 8  # p = ceil((32 + ceil(log2 T)) / 8)
    return {
        "blend_fwd": (28 + 4 * C) * K + (4 * C + 8) * P + 8 * T,
        "blend_bwd": (28 + 4 * C) * K + (4 * C + 8) * P + 2 * (24 + 4 * C) * N,
        "project_fwd": 24 * N, "project_bwd": 36 * N, "cov3d_fwd": 53 * N, "cov3d_bwd": 81 * N,
        "ewa_fwd": 65 * N, "ewa_bwd": 93 * N,
        "sort": 28 * N + 28 * K + 24 * key_bytes * K + 8 * T,  # global-radix model; a segmented design is credited the same
    }


# ----------------------------------------------------------------------------- clocks
class Clocks:
    """SM clock and throttle reasons (NVML), sampled from the MAIN thread while a block's kernels are in flight: no
    sampler thread competes with the launching thread for the core (round 1's 2 ms poller perturbed the steps)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.nv = [], set(), None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        if nv is None:
            return
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "how": "NVML from the launching thread while each timed block is in flight"}


# ----------------------------------------------------------------------------- reference arm (CPU port)
def cpu_reference(N, W, H, profile, steps, warmup, budget_s):
    """Times the CPU port of the path (oracle/splat_oracle.c) on all host cores.

    steps=None: as many full render steps as fit in ~budget_s (cpu_baseline leg).
    steps=K:    exactly K steps; if K full frames would exceed ~200 s each step renders only the
                top rows of the frame (a bounded sample) and the rate is scaled by the row fraction.
    """
    syn = synthetic_module()
    from oracle import c_oracle as C

    sc = syn.make_scene(N, W, H, seed=0, profile=profile)
    cores = C.num_threads()

    def make_step(Hs):
        Gimg = syn.make_grad_image(3, W, Hs)
        intr = sc.intr.clone()
        return lambda: C.render_step_fwd_bwd(sc.xyz, sc.scale, sc.rotate, sc.opacity, sc.rgb, intr, sc.extr, sc.bg, W,
                                             Hs, Gimg)

    step = make_step(H)
    t0 = time.perf_counter()
    _, _, info = step()
    t_one = time.perf_counter() - t0
    Hs = H
    if steps is None:
        steps = max(3, min(100, int(budget_s / max(t_one, 1e-3))))
    elif steps * t_one > 200.0:
        Hs = max(16, int(H * 200.0 / (steps * t_one)) // 16 * 16)
        step = make_step(Hs)
        _, _, info = step()
    for _ in range(max(0, warmup - 1)):
        if time.perf_counter() - t0 > 30.0:
            break
        step()
    t1 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t1
    frac = Hs / H
    what = "full render steps" if Hs == H else f"render steps on the top {Hs} of {H} rows (rate scaled by {frac:.3f})"
    return {"value": frac * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} {what} ({N} Gaussians, {W}x{H}, K={info['K']}) of oracle/splat_oracle.c "
                      f"(OpenMP, {cores} threads)", "ms_per_step": 1e3 * dt / steps / frac, "steps": steps, "K": info["K"]}


def workload_string(args, N, W, H):
    return f"{args.workload}: {N} Gaussians, {W}x{H}, render step fwd+bwd (C=3), profile {args.profile}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, W, H = WORKLOADS[args.workload]
    base = cpu_reference(N, W, H, args.profile, args.steps, args.warmup, budget_s=20.0)
    sample = base["sample"]
    if args.gpus > 1:
        sample += f"; one host: the CPU arm does not shard, the same figure is printed for --gpus {args.gpus}"
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": base["steps"], "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args, N, W, H), "K": base["K"]},
        "cpu_baseline": {"value": base["value"], "unit": UNIT, "cores": base["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def pin_to_own_cores(local_rank, world):
    """Each rank of a node gets its own slice of the cores the job may run on: eight ranks on one 32-core affinity set
    otherwise migrate over each other's cores and the slowest rank's step stretches."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(world, 1)
        if world > 1 and per >= 2:
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            return per
        return len(cores)
    except (AttributeError, OSError):
        return None


def run_ours(args):
    # stdout carries exactly ONE JSON line.  NCCL_DEBUG=INFO / VERSION output is written to fd 1 by NCCL itself: send
    # fd 1 to stderr for the whole run (the driver still sees the NCCL log there) and keep the real stdout for the line.
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores_per_rank = pin_to_own_cores(local_rank, world)

    import gflow_b200 as G
    from gflow_b200 import capi, frames, hostapi
    from gflow_b200.synthetic import make_camera, make_grad_image, make_scene

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    t_init = t_bcast = t_bcast_first = t_bcast_frame = 0.0
    if distributed:
        t0 = time.perf_counter()
        dist.init_process_group("nccl", device_id=dev)
        # communicator bring-up (hundreds of ms) happens here, not inside the first payload collective
        frames.warm_up(dev)
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t0
    lib = capi.load()
    N, W, H = WORKLOADS[args.workload]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T, P = gx * gy, W * H

    # ---- state: rank 0 builds the Gaussians, one NCCL broadcast hands them to every rank (8e);
    #      each rank then works on its own frame (own camera + own target gradient image).
    sc = make_scene(N, W, H, seed=0, profile=args.profile)
    state = {k: getattr(sc, k).to(dev) for k in ("xyz", "scale", "rotate", "opacity", "rgb")}
    if distributed:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        state = frames.broadcast_state(state if rank == 0 else None, src=0, device=dev)
        torch.cuda.synchronize()
        t_bcast_first = time.perf_counter() - t0  # first call at this size: includes the allocator growing on every rank
        dist.barrier()
        t0 = time.perf_counter()
        state = frames.broadcast_state(state if rank == 0 else None, src=0, device=dev)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - t0
    gen = torch.Generator().manual_seed(1000 + rank)
    intr, extr = make_camera(W, H, gen) if rank > 0 else (sc.intr, sc.extr)
    intr, extr = intr.to(dev), extr.to(dev)
    Gimg = make_grad_image(3, W, H, seed=1 + rank).to(dev)
    params = [state[k].clone().requires_grad_(True) for k in ("xyz", "scale", "rotate", "opacity", "rgb")]
    extr_p = extr.clone().requires_grad_(True)
    use_sh = args.workload == "cfg5"  # BASELINE config 5: colour from degree-3 spherical harmonics
    cam_center = -(extr[:, :3].T @ extr[:, 3])
    if use_sh:
        gsh = torch.Generator().manual_seed(7)
        shs = (torch.randn(N, 3, 16, generator=gsh) * 0.2).to(dev).requires_grad_(True)
        params[4] = shs

    def sh_colour(feature, xyz):
        return (G.compute_sh(feature, xyz - cam_center) + 0.5).clamp_min(0.0)

    def step(raster=G.rasterization):
        for p in params:
            p.grad = None
        extr_p.grad = None
        xyz, scale, rot, op, rgb = params
        if use_sh:
            rgb = sh_colour(rgb, xyz)
        img = raster(xyz, scale, rot, op, rgb, intr, extr_p, W, H, sc.bg)
        # loss = sum(out * G) with a fixed random G: dL/dout = G is fed to autograd directly
        img.backward(Gimg)
        return img

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        if not args.no_flush:
            flush_buf.fill_(1)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local_rank)

    def timed_blocks(fn, steps, blocks, sample_clocks=False):
        """`blocks` blocks of `steps` calls of fn(); every block bracketed by barrier + synchronize, CUDA events around
        each call, L2 flushed between calls.  Returns (median over blocks of the max-over-ranks block time [ms],
        per-block times, all per-step times of this rank, wall seconds)."""
        per_block, all_steps = [], []
        wall = 0.0
        for _ in range(blocks):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            t0 = time.perf_counter()
            for a, b in ev:
                flush_l2()
                a.record()
                fn()
                b.record()
            if sample_clocks:
                clocks.sample()  # the block's kernels are still in flight
            barrier()
            wall += time.perf_counter() - t0
            ms = [a.elapsed_time(b) for a, b in ev]
            all_steps += ms
            t_block = sum(ms)
            if distributed:
                tt = torch.tensor([t_block], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_block = float(tt.item())
            per_block.append(t_block)
        return statistics.median(per_block), per_block, all_steps, wall

    # ---- value: device-resident inputs
    W_UP = max(3, args.warmup)
    for _ in range(W_UP):
        step()
    launches0 = lib.gfb_kernel_launch_count()
    blk_ms, blocks_ms, step_ms, t_wall = timed_blocks(step, args.steps, max(1, args.blocks), sample_clocks=True)
    launches = (lib.gfb_kernel_launch_count() - launches0) // max(1, args.blocks)
    value = world * args.steps / (blk_ms / 1e3)

    # ---- the same step as ONE CUDA graph (gflow_b200.GraphedRenderStep): no autograd engine, one launch per step
    graphed = None
    if not use_sh:
        gstep = G.GraphedRenderStep(*[p.detach() for p in params], intr, extr, W, H, sc.bg)
        gstep.g_image.copy_(Gimg)
        for _ in range(3):
            gstep()
        g_ms, g_blocks, _, _ = timed_blocks(gstep, args.steps, max(1, args.blocks))
        gstep.check()
        graph_kernels = gstep.kernels_per_step
        graphed = {"value": world * args.steps / (g_ms / 1e3), "unit": UNIT, "ms_per_step": g_ms / args.steps,
                   "blocks_ms": [round(b, 4) for b in g_blocks], "K": gstep.k(), "capacity": gstep.capacity,
                   "what": "gflow_b200.GraphedRenderStep: gfb_render_forward + gfb_render_backward captured once into a CUDA "
                           "graph over static buffers, one cudaGraphLaunch per step (same kernels, same inputs, same L2 "
                           "flush); the eager autograd step above spends ~125 us of host time per step"}
        # independent frames side by side: FRAMES_IN_FLIGHT graphed steps (own buffers, own stream each), what a rank of
        # the frame-sharded fit does with its frames; one sequential step leaves SM time unused in every kernel's tail
        others = [G.GraphedRenderStep(*[p.detach() for p in params], intr, extr, W, H, sc.bg, capacity=gstep.capacity)
                  for _ in range(FRAMES_IN_FLIGHT - 1)]
        for o_ in others:
            o_.g_image.copy_(Gimg)
        gsteps = [gstep] + others
        side = [torch.cuda.Stream(device=dev) for _ in gsteps]
        per_call = 5

        def frames_in_flight():
            cur = torch.cuda.current_stream(dev)
            for s_ in side:
                s_.wait_stream(cur)
            for _ in range(per_call):
                for g_, s_ in zip(gsteps, side):
                    with torch.cuda.stream(s_):
                        g_.graph.replay()
            for s_ in side:
                cur.wait_stream(s_)

        frames_in_flight()
        n_calls = max(2, args.steps // (per_call * FRAMES_IN_FLIGHT))
        c_ms, _, _, _ = timed_blocks(frames_in_flight, n_calls, max(1, min(args.blocks, 5)))
        for g_ in gsteps:
            g_.check()
        n_steps = n_calls * per_call * FRAMES_IN_FLIGHT
        graphed["frames_in_flight"] = {
            "frames": FRAMES_IN_FLIGHT, "value": world * n_steps / (c_ms / 1e3), "unit": UNIT, "ms_per_step": c_ms / n_steps,
            "what": f"{FRAMES_IN_FLIGHT} independent graphed steps (own buffers) replayed round-robin on {FRAMES_IN_FLIGHT} "
                    "streams: aggregate steps/s of one GPU when frames do not depend on each other"}
        del others, gsteps
        # a batch of cameras over one set of Gaussians (BASELINE config 5's "8-frame batch"): gflow_b200.BatchedRenderStep,
        # frames side by side on streams, joined per batch, per-frame gradients summed by one kernel
        batch = {}
        for n_f in (4, 8):
            cams = [make_camera(W, H, torch.Generator().manual_seed(2000 + 17 * rank + f)) for f in range(n_f)]
            bstep = G.BatchedRenderStep(*[p.detach() for p in params], torch.stack([c[0] for c in cams]).to(dev),
                                        torch.stack([c[1] for c in cams]).to(dev), W, H, sc.bg)
            for g_ in bstep.g_images:
                g_.copy_(Gimg)
            for _ in range(3):
                bstep()
            n_b = max(3, args.steps // n_f)
            b_ms, _, _, _ = timed_blocks(bstep, n_b, max(1, min(args.blocks, 5)))
            bstep.check()
            batch[f"frames_{n_f}"] = {"value": world * n_b * n_f / (b_ms / 1e3), "unit": UNIT, "ms_per_batch": b_ms / n_b}
            del bstep
        batch["what"] = ("gflow_b200.BatchedRenderStep: F cameras over one set of Gaussians, forward + backward side by side on F "
                         "streams, joined per batch, parameter gradients summed over the views; frames/s")
        graphed["batch_of_frames"] = batch
        del gstep

    # ---- e2e: the same step with pinned HOST buffers (gflow_b200.hostapi.HostRenderStep): H2D of the step's inputs
    #      and D2H of loss + gradients inside the timed region, copies double buffered against the kernels
    host_in = torch.empty(sum(p.numel() for p in params) + 16, dtype=torch.float32).pin_memory()
    o = 0
    for t in [p.detach() for p in params] + [intr, extr]:
        host_in[o:o + t.numel()].copy_(t.reshape(-1).cpu())
        o += t.numel()
    host = hostapi.HostRenderStep(N, W, H, tuple(params[4].shape[1:]), Gimg, sc.bg, dev, depth=args.e2e_depth,
                                  colour=sh_colour if use_sh else None, sample_input=host_in)
    n_slot = max(1, args.e2e_depth)
    host_outs = [host.host_output_block() for _ in range(n_slot)]
    n_e2e = max(5, min(args.steps, 50))
    for i in range(3):
        host.submit(host_in, host_outs[i % n_slot])
    host.wait()
    e2e_blocks = []
    for _ in range(max(1, min(args.blocks, 5))):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush_l2()
        a.record()
        for i in range(n_e2e):
            host.submit(host_in, host_outs[i % n_slot])
        host.wait()          # every D2H has landed
        b.record()
        barrier()
        t = a.elapsed_time(b)
        if distributed:
            tt = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        e2e_blocks.append(t)
    host.check()
    e2e_value = world * n_e2e / (statistics.median(e2e_blocks) / 1e3)
    # the same copies with no kernels between them (both directions at once, as in the step): the link's own ceiling
    cur = torch.cuda.current_stream(dev)
    link_ms = []
    for _ in range(3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        host.h2d_stream.wait_stream(cur)
        host.d2h_stream.wait_stream(cur)
        for i in range(n_e2e):
            with torch.cuda.stream(host.h2d_stream):
                host.slots[i % n_slot]["dev_in"].copy_(host_in, non_blocking=True)
            with torch.cuda.stream(host.d2h_stream):
                host_outs[i % n_slot].copy_(host.slots[i % n_slot]["dev_out"], non_blocking=True)
        cur.wait_stream(host.h2d_stream)
        cur.wait_stream(host.d2h_stream)
        b.record()
        barrier()
        link_ms.append(a.elapsed_time(b) / n_e2e)
    link_ms = statistics.median(link_ms)
    e2e_loss = float(host.unpack_output(host_outs[(n_e2e - 1) % n_slot])["loss"][0])

    # ---- roofline of the dominant kernel (alpha-blending backward), timed alone with CUDA events
    roof = kernel_roofline(G, lib, params, intr, extr, Gimg, sc.bg, N, W, H, T, P, flush_l2, dev, clocks,
                           cam_center if use_sh else None, args.profile)

    if use_sh:
        roof["compute_sh"] = sh_roofline(G, params[4].detach(), (params[0].detach() - cam_center), flush_l2)

    chain = iteration = sequence = None
    if not args.quick:
        # ---- the same step through the five separate operators, exactly as render.py:21-64 calls them
        for _ in range(3):
            step(G.rasterization_unfused)
        c_ms, _, _, _ = timed_blocks(lambda: step(G.rasterization_unfused), max(5, min(args.steps, 50)), 3)
        n_chain = max(5, min(args.steps, 50))
        chain = {"value": world * n_chain / (c_ms / 1e3), "unit": UNIT, "ms_per_step": c_ms / n_chain,
                 "what": "same step through project_point/compute_cov3d/ewa_project/sort_gaussian/alpha_blending "
                         "called one by one (render.py:21-64 pattern)"}
        # ---- SURVEY 8d unit (iii): one GFlow iteration's rendering work
        if not use_sh:
            it_fn = make_gflow_iteration(G, params, intr, extr_p, W, H, sc.bg, dev)
            for _ in range(3):
                it_fn()
            n_it = max(5, min(args.steps, 50))
            i_ms, _, _, _ = timed_blocks(it_fn, n_it, 3)
            iteration = {"value": world * n_it / (i_ms / 1e3), "unit": "iters/s", "ms_per_step": i_ms / n_it,
                         "what": "render_multiple call pattern through `import msplat` (project, cov3d, ewa, sort, four "
                                 "blends: rgb C=3, depth C=1, depth-colour C=3, centre C=3) + backward of an rgb and a "
                                 "depth loss (render.py:6-108, trainer.py:404-533)"}

    # ---- end of sequence: one gather of per-frame outputs (rendered frame + pose) on rank 0
    t_gather = 0.0
    if distributed:
        with torch.no_grad():
            img = G.rasterization(*[p.detach() for p in params[:4]],
                                  sh_colour(params[4].detach(), params[0].detach()) if use_sh else params[4].detach(),
                                  intr, extr, W, H, sc.bg)
        barrier()
        t0 = time.perf_counter()
        frames.gather_frames(img, extr, dst=0)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - t0
        # the checkpoint-shaped frame state (attributes + camera + still mask + last_uv) over NCCL as well (8f rank 4)
        t_bcast_frame = broadcast_frame_state_roundtrip(frames, state, intr, extr, W, H, dev, rank)

    # ---- BASELINE config 4: frame-sharded sequence fit through the native loop
    if not args.quick and not args.no_sequence and args.workload == "cfg2":
        sequence = sequence_section(dist if distributed else None, world, rank, dev)

    # ---- BASELINE config 3 (300-iteration per-frame Adam loop), operator path and native path, each in a
    #      process of its own (tools/bench_fit.py) so a fault there cannot touch the numbers above
    fit_loop = proxy = cpu = None
    ref_gpu = None
    if not use_sh:
        ref_gpu = reference_gpu_section(params, intr, extr_p, W, H, sc.bg, Gimg, timed_blocks, args.steps, world)
    single = rank == 0 and world == 1
    if single and args.workload == "cfg2" and not args.no_fit_loop and not args.quick:
        fit_loop = fit_loop_section()
    if single and not args.no_proxy and not args.quick and args.workload == "cfg2":
        proxy = eager_proxy_section(args)
    if single and not args.no_cpu_baseline:
        cpu = cpu_reference(N, W, H, args.profile, None, 1, budget_s=12.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    # `value`: the render step with inputs resident in HBM.  Where the step can be one CUDA graph (plain colours) that is
    # the figure -- same kernels, and the GPU is the only limit; the eager autograd call of the same step is host-bound
    # (107 us of host per 106 us of kernels) and follows the box's CPU (4 800 - 9 300 iters/s over this round's boxes),
    # it is reported beside it as `eager`.  With an SH colour callback (cfg5) the eager step is the step.
    eager = {"value": value, "unit": UNIT, "ms_per_step": blk_ms / args.steps, "blocks_ms": [round(b, 4) for b in blocks_ms],
             "gpu_launches": int(launches), "ms_per_step_median": statistics.median(step_ms),
             "api": f"msplat.rasterization (gflow_b200.ops, fused pipeline, {G.BACKEND} binding) -> C ABI, eager autograd"}
    value_api = eager["api"]
    if graphed is not None:
        value, blk_ms, blocks_ms = graphed["value"], graphed["ms_per_step"] * args.steps, graphed["blocks_ms"]
        launches = graph_kernels * args.steps
        value_api = ("gflow_b200.GraphedRenderStep: gfb_render_forward_keep + gfb_render_backward_keep (C ABI) captured once, "
                     "one cudaGraphLaunch per step")
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W_UP, "ms_per_step": blk_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args, N, W, H),
                       "K": roof.pop("K"), "sharding": "one frame (camera + target) per rank, no in-loop collective",
                       "l2": "256 MiB written between timed steps" if not args.no_flush else "not flushed (working set < L2)",
                       "api": value_api,
                       "timing": f"median of {max(1, args.blocks)} blocks of {args.steps} steps, max over ranks per block"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes, "d2h_bytes_per_step": host.d2h_bytes,
                    "steps": n_e2e, "copies_alone_ms_per_step": link_ms,
                    "copies_alone_GBps_per_direction": max(host.h2d_bytes, host.d2h_bytes) / (link_ms * 1e6),
                    "steps_in_flight": n_slot,
                    "api": f"gflow_b200.hostapi.HostRenderStep (pinned host blocks, one copy stream per direction, {n_slot} "
                           "independent steps in flight, compute as "
                           + ("one CUDA graph per slot on the slot's own stream, one library call per step)" if host.graphed
                              else "eager autograd: SH colour)"),
                    "loss_read_back": e2e_loss},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roof,
            "blocks_ms": [round(b, 4) for b in blocks_ms],
            "eager": eager,
            "ms_per_step_median": statistics.median(step_ms) if graphed is None else blk_ms / args.steps,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / (args.steps * max(1, args.blocks)),
        }
        if cores_per_rank is not None:
            line["config"]["host_cores_per_rank"] = cores_per_rank
        for k, v in (("graphed", graphed), ("operator_chain", chain), ("gflow_iteration", iteration), ("sequence", sequence), ("cpu_baseline", cpu),
                     ("fit_loop", fit_loop), ("gpu_proxy_baseline", proxy), ("reference_gpu", ref_gpu)):
            if v is not None:
                line[k] = v
        if distributed:
            line["collectives_ms"] = {"nccl_init_and_warmup": 1e3 * t_init, "broadcast_state": 1e3 * t_bcast,
                                      "broadcast_state_first_call": 1e3 * t_bcast_first,
                                      "gather_frames": 1e3 * t_gather, "broadcast_frame_state": 1e3 * t_bcast_frame}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if distributed:
        dist.destroy_process_group()


def broadcast_frame_state_roundtrip(frames, state, intr, extr, W, H, dev, rank):
    """frames.broadcast_frame_state (the packed checkpoint wire format of gflow_b200/checkpoint.py) over NCCL, checked
    on every rank against the state it already holds.  Returns the seconds of the (warm) broadcast."""
    from gflow_b200 import checkpoint

    n = state["xyz"].shape[0]
    fs = checkpoint.FrameState(attributes={k: v.detach() for k, v in state.items()}, intr=intr, extr=extr,
                               still_mask=torch.arange(n, device=dev) % 3 == 0,
                               last_uv=state["xyz"][:, :2].detach().contiguous(), width=W, height=H)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = frames.broadcast_frame_state(fs if rank == 0 else None, src=0, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for k in state:
        if not torch.equal(got.attributes[k], state[k]):
            raise SystemExit(f"bench.py: broadcast_frame_state delivered a different {k} on rank {rank}")
    if not torch.equal(got.still_mask.to(dev), fs.still_mask) or int(got.width) != W:
        raise SystemExit(f"bench.py: broadcast_frame_state delivered a different still mask / size on rank {rank}")
    return dt


def make_gflow_iteration(G, params, intr, extr_p, W, H, bg, dev):
    """The rendering work of one GFlow iteration, called exactly as render_multiple calls it (render.py:6-108) through
    the drop-in module name; the depth colour map stays on the device (the reference's matplotlib round trip is host
    work of the caller, not of the operators).  Loss: mean rgb + 0.1 mean depth (trainer.py:452-488 in shape)."""
    G.install_dropin()
    import msplat

    ident = torch.tensor([1.0, 0.0, 1.0], device=dev)

    def iteration():
        for p in params:
            p.grad = None
        extr_p.grad = None
        xyz, scale, rot, op, rgb = params
        uv, depth = msplat.project_point(xyz, intr, extr_p, W, H)
        visible = depth != 0
        cov3d = msplat.compute_cov3d(scale, rot, visible)
        conic, radius, tiles = msplat.ewa_project(xyz, cov3d, intr, extr_p, uv, W, H, visible)
        ids, rng = msplat.sort_gaussian(uv, depth, W, H, radius, tiles)
        r_rgb = msplat.alpha_blending(uv, conic, op, rgb, ids, rng, bg, W, H)
        r_depth = msplat.alpha_blending(uv, conic, op, depth, ids, rng, bg, W, H)
        with torch.no_grad():
            dcol = (depth.detach() * 0.2).clamp(0, 1).expand(-1, 3).contiguous()
            msplat.alpha_blending(uv, conic, op, dcol, ids, rng, bg, W, H)
            msplat.alpha_blending(uv, torch.ones_like(conic) * ident, torch.ones_like(op), rgb, ids, rng, bg, W, H)
        (r_rgb.mean() + 0.1 * r_depth.mean()).backward()

    return iteration


def sequence_section(dist, world, rank, dev):
    """BASELINE config 4: SEQ_FRAMES synthetic frames (60k Gaussians, 854x480), SEQ_ITERS native Adam iterations each
    (mse + depth loss), frames sharded across the ranks; wall clock from before the state broadcast to after the gather
    of the per-frame outputs, max over ranks."""
    from gflow_b200 import fit
    from gflow_b200.synthetic import make_camera, make_scene

    N, W, H = WORKLOADS["cfg2"]
    sc = make_scene(N, W, H, seed=0, profile="gflow")
    raw = {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
           "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
           "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}
    raw_dev = {k: v.to(dev) for k, v in raw.items()}
    pose0 = fit.extr_to_pose(sc.extr)

    class Targets:
        """Per-frame synthetic priors, produced locally by the rank that owns the frame (the reference reads them from
        disk): the scene rendered from a slightly different camera, plus its depth map."""

        def __len__(self):
            return SEQ_FRAMES

        def __call__(self, i):
            gen = torch.Generator().manual_seed(100 + i)
            _, extr = make_camera(W, H, gen)
            f = fit.FrameFitter(raw_dev, sc.intr.to(dev), fit.extr_to_pose(extr).to(dev), W, H)
            with torch.no_grad():
                img, dmap, _ = f.render(0.0, want_depth=True)
            return img.permute(1, 2, 0).contiguous(), dmap.permute(1, 2, 0).contiguous(), pose0

    cfg = fit.FitConfig(iterations=SEQ_ITERS, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True)
    targets = Targets()
    warm = fit.FitConfig(iterations=5, lr=4e-3, lr_camera=1e-3, lambda_depth=0.1, native=True)
    fit.FrameFitter(raw_dev, sc.intr.to(dev), pose0.to(dev), W, H).train(*targets(0)[:2], warm)
    runs = []
    for _ in range(SEQ_RUNS):  # the whole sequence SEQ_RUNS times, median reported: its host side (targets, per-frame set-up,
        if dist is not None:   # status reads) makes one run sensitive to a busy host
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        results, gathered = fit.fit_sequence_sharded(raw_dev if rank == 0 or dist is None else None, sc.intr, targets, W, H, cfg,
                                                     dev, concurrent_frames=SEQ_CONCURRENT)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        runs.append(dt)
    dt = statistics.median(runs)
    first = results[min(results)]
    per_rank = (SEQ_FRAMES + world - 1) // world
    return {"what": f"BASELINE config 4: {SEQ_FRAMES} synthetic frames x {SEQ_ITERS} native Adam iterations (60k Gaussians, "
                    "854x480, mse + depth loss), frames sharded across the ranks, each rank working on "
                    f"{SEQ_CONCURRENT} of its (independent) frames at a time on separate streams; wall clock incl. the state "
                    "broadcast and the gather of per-frame outputs, max over ranks",
            "value": SEQ_FRAMES * SEQ_ITERS / dt, "unit": "frame-iterations/s", "seconds": dt, "frames": SEQ_FRAMES,
            "iterations_per_frame": SEQ_ITERS, "frames_per_rank": per_rank, "n_gpus": world,
            "concurrent_frames_per_gpu": SEQ_CONCURRENT, "runs_seconds": [round(r, 4) for r in runs],
            "loss_first": first.losses[0], "loss_last": first.losses[-1]}


def fit_loop_section():
    """BASELINE config 3: 60k Gaussians, 854x480, per-frame Adam loop with rgb + depth loss, iterations/s.
    `operator_path` = msplat operators + autograd + torch.optim.Adam (how gflow/trainer.py drives them);
    `native` = the same iteration as eight kernels (csrc/fit.cu); `native_ssim` adds the 1 - SSIM term GFlow's
    loss_rgb carries (trainer.py:459-462)."""
    import subprocess

    out = {"what": "BASELINE config 3: 60k Gaussians, 854x480, per-frame Adam loop (mse + depth loss), iterations/s",
           "unit": "iters/s"}
    tool = os.path.join(ROOT, "tools", "bench_fit.py")
    for key, extra, iters in (("operator_path", [], 100), ("native", ["--native"], 300),
                              ("native_ssim", ["--native", "--ssim"], 300)):
        try:
            res = subprocess.run([sys.executable, tool, "--iters", str(iters), *extra], capture_output=True, text=True,
                                 timeout=120)
            rec = None
            for ln in res.stdout.splitlines():
                if ln.startswith("{"):
                    rec = json.loads(ln)
            if rec is None:
                out[key] = {"error": (res.stderr or res.stdout)[-300:]}
            else:
                out[key] = {"value": rec["value"], "iterations": rec["iterations"], "seconds": rec["seconds"],
                            "loss_first": rec["loss_first"], "loss_last": rec["loss_last"]}
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": repr(e)[:300]}
    return out


def find_real_msplat():
    """A real MSplat build (github.com/pointrix-project/msplat), should one be on the box: `baseline/_ref` first, then
    site-packages -- never this repository's drop-in.  Loaded under an alias so it coexists with the drop-in `msplat`.
    Returns (module or None, list of probed directories)."""
    probed = [os.path.join(ROOT, "baseline", "_ref")] + [p for p in sys.path if "site-packages" in p]
    for base in probed:
        init = os.path.join(base, "msplat", "__init__.py")
        try:
            if os.path.exists(init) and "gflow_b200" not in open(init).read():
                spec = importlib.util.spec_from_file_location("_real_msplat", init,
                                                              submodule_search_locations=[os.path.dirname(init)])
                mod = importlib.util.module_from_spec(spec)
                sys.modules["_real_msplat"] = mod
                spec.loader.exec_module(mod)
                return mod, probed
        except Exception:  # noqa: BLE001  (a broken install is reported as absent, with the probe list)
            pass
    return None, probed


def reference_gpu_section(params, intr, extr_p, W, H, bg, Gimg, timed_blocks, steps, world):
    """SURVEY.md 8d "Reference GPU path": the identical render step (five operators, C = 3 blend, full backward) through a
    real MSplat on this GPU, in the same process and harness -- when one can be imported.  Otherwise says so, with what
    was probed; the eager-PyTorch proxy (`gpu_proxy_baseline`) is then the only GPU-side comparator, and it is not MSplat."""
    real, probed = find_real_msplat()
    if real is None:
        return {"available": False, "probed": probed,
                "note": "no MSplat build on this box (not vendored by the reference, unpinned, no network): the target "
                        "'>= 1.5x MSplat' cannot be measured here; tests/test_gpu_parity.py::test_against_real_msplat "
                        "makes the same probe and records a golden set on first contact"}

    def step():
        for p in params:
            p.grad = None
        extr_p.grad = None
        xyz, scale, rot, op, rgb = params
        uv, depth = real.project_point(xyz, intr, extr_p, W, H)
        vis = depth != 0
        cov = real.compute_cov3d(scale, rot, vis)
        conic, radius, tiles = real.ewa_project(xyz, cov, intr, extr_p, uv, W, H, vis)
        ids, rng = real.sort_gaussian(uv, depth, W, H, radius, tiles)
        real.alpha_blending(uv, conic, op, rgb, ids, rng, bg, W, H).backward(Gimg)

    for _ in range(5):
        step()
    n = max(5, min(steps, 50))
    ms, _, _, _ = timed_blocks(step, n, 3)
    return {"available": True, "module": getattr(real, "__file__", "?"), "value": world * n / (ms / 1e3), "unit": UNIT,
            "ms_per_step": ms / n, "what": "the operator-chain render step through the real MSplat, same inputs and harness"}


def eager_proxy_section(args):
    """with MSplat unavailable, the only GPU-side comparator the plan allows is the
    oracle's differentiable PyTorch restatement executed with CUDA tensors (eager PyTorch).  A labelled PROXY, not
    MSplat; run in a process of its own (tools/eager_proxy.py), bounded to a few steps."""
    import subprocess

    tool = os.path.join(ROOT, "tools", "eager_proxy.py")
    try:
        res = subprocess.run([sys.executable, tool, args.workload, args.profile, "2"], capture_output=True, text=True, timeout=240)
        for ln in res.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (res.stderr or res.stdout)[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def sh_roofline(G, shs, dirs, flush_l2, reps=20):
    """compute_sh forward / backward timed alone: the one purely streaming kernel of the path ((4CK + 12 + 4C) N bytes
    forward, (8CK + 24 + 4C) N backward), against the measured HBM copy bandwidth."""
    peak, _ = load_peaks()
    N, C, K = shs.shape
    s = shs.clone().requires_grad_(True)
    d = dirs.clone().requires_grad_(True)
    g = torch.randn(N, C, device=shs.device)

    def timeit(fn):
        ts = []
        for _ in range(reps + 3):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts[3:])

    with torch.no_grad():
        ms_f = timeit(lambda: G.compute_sh(s, d))
    out = G.compute_sh(s, d)

    def bwd():
        s.grad = d.grad = None
        out.backward(g, retain_graph=True)

    ms_b = timeit(bwd)  # includes the autograd engine's launch path; the kernel itself is the only GPU work
    bf, bb = (4 * C * K + 12 + 4 * C) * N, (8 * C * K + 24 + 4 * C) * N
    # what a plain device copy moving the SAME number of bytes takes under the same harness: at tens of MB the launch
    # and DRAM ramp are a visible part of any kernel, so this is the practical ceiling next to the 2 GiB-copy peak
    src_f, src_b = torch.empty(bf // 8, device=shs.device), torch.empty(bb // 8, device=shs.device)
    dst_f, dst_b = torch.empty_like(src_f), torch.empty_like(src_b)
    cp_f, cp_b = timeit(lambda: dst_f.copy_(src_f)), timeit(lambda: dst_b.copy_(src_b))
    return {"forward": {"kernel_ms": ms_f, "algorithmic_bytes": bf, "achieved": bf / (ms_f * 1e-3) / 1e9, "frac": bf / (ms_f * 1e-3) / 1e9 / peak,
                        "same_bytes_copy_ms": cp_f, "frac_of_same_bytes_copy": cp_f / ms_f},
            "backward": {"kernel_ms": ms_b, "algorithmic_bytes": bb, "achieved": bb / (ms_b * 1e-3) / 1e9, "frac": bb / (ms_b * 1e-3) / 1e9 / peak,
                         "same_bytes_copy_ms": cp_b, "frac_of_same_bytes_copy": cp_b / ms_b},
            "unit": "GB/s", "what": f"compute_sh degree-{int(K ** 0.5) - 1} ({N} Gaussians x {C} x {K} coefficients), CUDA events around the op call, L2 flushed"}


def ncu_constants(N, W, H, profile):
    """Per-launch constants of the alpha-blending backward from the committed ncu --set full capture of this workload
    (profiles/traffic.json): DRAM bytes and executed warp instructions.  None when no capture matches."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tp) as fh:
            rec = json.load(fh)
        for r in rec.get("captures", []):
            if (r["N"], r["W"], r["H"], r["profile"]) == (N, W, H, profile) and "path" not in r:  # operator-level capture
                return r
    except Exception:
        pass
    return None


def kernel_roofline(G, lib, params, intr, extr, Gimg, bg, N, W, H, T, P, flush_l2, dev, clocks, cam_center, profile, reps=30):
    """Average duration of gfb_alpha_blending_bwd alone (CUDA events on the launching stream)."""
    from gflow_b200 import capi

    peak, peak_src = load_peaks()
    with torch.no_grad():
        xyz, scale, rot, op, rgb = [p.detach() for p in params]
        if rgb.dim() == 3:
            rgb = (G.compute_sh(rgb, xyz - cam_center) + 0.5).clamp_min(0.0).contiguous()
        uv, depth = G.project_point(xyz, intr, extr, W, H)
        vis = depth != 0
        cov = G.compute_cov3d(scale, rot, vis)
        conic, radius, tiles = G.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
        ids, rng = G.sort_gaussian(uv, depth, W, H, radius, tiles)
        K = ids.numel()
        st = torch.cuda.current_stream().cuda_stream
        geom = torch.empty(max(K, 1) * 8, device=dev)
        feat = torch.empty(max(K, 1) * 4, device=dev)
        out = torch.empty(3, H, W, device=dev)
        fT = torch.empty(H, W, device=dev)
        nc = torch.empty(H, W, device=dev, dtype=torch.int32)
        gp = torch.zeros(N * 12, device=dev)
        opf = op.reshape(-1).contiguous()
        capi.check(lib.gfb_blend_pack_geometry(uv.data_ptr(), conic.data_ptr(), opf.data_ptr(), ids.data_ptr(), K,
                                               geom.data_ptr(), st), "pack geometry")
        capi.check(lib.gfb_blend_pack_feature(rgb.data_ptr(), 3, 0, 3, ids.data_ptr(), K, feat.data_ptr(), st), "pack feature")

        def fwd():
            capi.check(lib.gfb_alpha_blending_fwd(geom.data_ptr(), feat.data_ptr(), K, rng.data_ptr(), 3, 0, 3, bg, W, H,
                                                  out.data_ptr(), fT.data_ptr(), nc.data_ptr(), st), "blend fwd")

        def bwd():
            capi.check(lib.gfb_alpha_blending_bwd(geom.data_ptr(), feat.data_ptr(), K, ids.data_ptr(), rng.data_ptr(), 3, 0,
                                                  3, bg, W, H, fT.data_ptr(), nc.data_ptr(), Gimg.data_ptr(),
                                                  gp.data_ptr(), st), "blend bwd")

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            return sum(ts) / len(ts)

        fwd()
        ms_f, ms_b = timeit(fwd), timeit(bwd)

        # SURVEY 8d unit (i): every other operator's kernels alone, same harness (C ABI calls, events, L2 flushed)
        visb = vis.reshape(-1).to(torch.uint8).contiguous()
        f32 = lambda *sh: torch.empty(*sh, device=dev)  # noqa: E731
        i32 = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.int32)  # noqa: E731
        o_uv, o_depth, o_cov, o_conic, o_rad, o_tiles = f32(N, 2), f32(N, 1), f32(N, 6), f32(N, 3), i32(N, 1), i32(N, 1)
        g_uv, g_depth, g_cov, g_conic = torch.rand(N, 2, device=dev), torch.rand(N, 1, device=dev), torch.rand(N, 6, device=dev), torch.rand(N, 3, device=dev)
        d_xyz, d_cam, d_scale, d_rot, d_cov = f32(N, 3), f32(16), f32(N, 3), f32(N, 4), f32(N, 6)
        xyz_c, scale_c, rot_c, intr_c, extr_c = (t.detach().contiguous() for t in (xyz, scale, rot, intr, extr))
        P_ = lambda t: t.data_ptr()  # noqa: E731
        tile_ws = torch.empty(lib.gfb_sort_tile_workspace_bytes(W, H), device=dev, dtype=torch.uint8)
        keys = torch.empty(max(K, 1), device=dev, dtype=torch.int64)
        ids2, rng2, k_host = i32(max(K, 1)), i32(T, 2), ctypes.c_int64(0)
        ops_alone = {
            "project_fwd": lambda: lib.gfb_project_point_fwd(P_(xyz_c), P_(intr_c), P_(extr_c), N, W, H, 0.2, 1.3, P_(o_uv), P_(o_depth), st),
            "project_bwd": lambda: lib.gfb_project_point_bwd(P_(xyz_c), P_(intr_c), P_(extr_c), N, W, H, 0.2, 1.3, P_(g_uv), P_(g_depth),
                                                             P_(d_xyz), P_(d_cam), st),
            "cov3d_fwd": lambda: lib.gfb_compute_cov3d_fwd(P_(scale_c), P_(rot_c), P_(visb), N, P_(o_cov), st),
            "cov3d_bwd": lambda: lib.gfb_compute_cov3d_bwd(P_(scale_c), P_(rot_c), P_(visb), N, P_(g_cov), P_(d_scale), P_(d_rot), st),
            "ewa_fwd": lambda: lib.gfb_ewa_project_fwd(P_(xyz_c), P_(cov), P_(intr_c), P_(extr_c), P_(uv), N, W, H, P_(visb), P_(o_conic),
                                                       P_(o_rad), P_(o_tiles), st),
            "ewa_bwd": lambda: lib.gfb_ewa_project_bwd(P_(xyz_c), P_(cov), P_(intr_c), P_(extr_c), P_(uv), N, W, H, P_(visb), P_(g_conic),
                                                       P_(d_xyz), P_(d_cov), P_(d_cam), st),
            "sort": lambda: lib.gfb_sort_gaussian(P_(uv), P_(depth), P_(radius), P_(tiles), N, W, H, P_(tile_ws), K, P_(keys), P_(ids2),
                                                  P_(rng2), ctypes.addressof(k_host), st),
        }
        ab_all = algorithmic_bytes(N, K, P, T)
        op_level = {}
        for name, fn_ in ops_alone.items():
            def checked(fn_=fn_, name=name):
                capi.check(fn_(), name)
            ms_ = timeit(checked)
            op_level[name] = {"kernel_ms": ms_, "algorithmic_bytes": ab_all[name], "achieved": ab_all[name] / (ms_ * 1e-3) / 1e9,
                              "frac": ab_all[name] / (ms_ * 1e-3) / 1e9 / peak}
    ab = algorithmic_bytes(N, K, P, T)
    ach = ab["blend_bwd"] / (ms_b * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "blend_bwd_kernel<3,false> (gfb_alpha_blending_bwd)", "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab["blend_bwd"], "kernel_ms": ms_b,
            "blend_fwd": {"kernel_ms": ms_f, "algorithmic_bytes_per_launch": ab["blend_fwd"],
                          "achieved": ab["blend_fwd"] / (ms_f * 1e-3) / 1e9, "frac": ab["blend_fwd"] / (ms_f * 1e-3) / 1e9 / peak},
            "note": "working set fits the 126 MB L2; the kernel is bound by the instruction issue rate, not by HBM (DESIGN.md)",
            "K": K,
            "op_level": dict(op_level, what="SURVEY 8d unit (i): each operator's kernels alone through the C ABI (GB/s of "
                                            "algorithmic bytes, frac of the HBM peak; sort = count+scan, scatter, per-tile "
                                            "sort incl. the wait for K; blend: blend_fwd above and the headline figures)")}
    cap = ncu_constants(N, W, H, profile)
    if cap is not None:
        mhz = clocks.summary().get("sm_mhz") or clocks.max_mhz or 1965
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        roof["traffic"] = cap["blend_bwd_dram_bytes_per_launch"]
        roof["traffic_source"] = cap["source"]
        # issue roofline: one warp instruction per scheduler per cycle, 4 schedulers per SM
        floor_ms = cap["blend_bwd_warp_instructions"] / (sms * 4 * mhz * 1e6) * 1e3
        roof["issue_frac"] = floor_ms / ms_b
        roof["issue_roofline"] = {"warp_instructions_per_launch": cap["blend_bwd_warp_instructions"], "sm_count": sms,
                                  "sm_mhz": mhz, "floor_ms": floor_ms,
                                  "what": "time to issue the kernel's warp instructions at 1 per scheduler per cycle / measured time"}
    return roof


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
