"""Kernel-logic tests that need no GPU: gflow_b200/csrc/*.cu compiled for the host against the SIMT shim
(tests/simt/) and driven through the C ABI with host pointers, compared with the CPU oracle.

What this checks: thread / tile indexing, warp collectives, CTA barriers, the double-buffered
cp.async.bulk + mbarrier staging protocol (poisoned until waited on), the segmented sort networks, the
scan, the gradient reductions.  What it cannot check: inline PTX itself, the memory model, FMA
contraction and MUFU rounding (hence image / gradient tolerances as on the GPU) -- the `-m gpu` tests
stay the parity tests of record.  The emulated library is test infrastructure, never a fallback.
"""
import os
import sys

import pytest
import torch

from conftest import assert_close
from gflow_b200.synthetic import make_grad_image, make_scene
from oracle import c_oracle as C

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import emu  # noqa: E402

IMG_OUTLIERS = dict(outlier_frac=2e-5, outlier_rel=2e-2)
GRAD_OUTLIERS = dict(outlier_frac=1e-4, outlier_rel=5e-2)


def _oracle(sc, Gimg, feature=None):
    feature = sc.rgb if feature is None else feature
    uv, depth = C.project_point(sc.xyz, sc.intr, sc.extr, sc.W, sc.H)
    vis = depth != 0
    cov = C.compute_cov3d(sc.scale, sc.rotate, vis)
    conic, radius, tiles = C.ewa_project(sc.xyz, cov, sc.intr, sc.extr, uv, sc.W, sc.H, vis)
    ids, rng = C.sort_gaussian(uv, depth, sc.W, sc.H, radius, tiles)
    img, grads, info = C.render_step_fwd_bwd(sc.xyz, sc.scale, sc.rotate, sc.opacity, feature, sc.intr, sc.extr, sc.bg,
                                             sc.W, sc.H, Gimg)
    return dict(uv=uv, depth=depth, cov3d=cov, conic=conic, radius=radius, tiles=tiles, ids=ids, tile_range=rng,
                image=img, grads=grads, K=info["K"])


def _check(r, o, what):
    assert r["K"] == o["K"], f"{what}: K {r['K']} vs oracle {o['K']}"
    for k in ("uv", "depth", "conic", "radius", "ids", "tile_range"):
        assert torch.equal(r[k], o[k]), f"{what}: {k} must be bit-exact"
    assert_close(r["image"], o["image"], 1e-4, f"{what} image", **IMG_OUTLIERS)
    for k in ("xyz", "scale", "rotate", "opacity", "feature", "intr", "extr"):
        assert_close(r["grads"][k].reshape(o["grads"][k].shape), o["grads"][k], 1e-3, f"{what} grad {k}",
                     **GRAD_OUTLIERS)


@pytest.fixture(params=[0, 1, 2], ids=["ascending", "descending", "shuffled"])
def schedule(request):
    emu.set_schedule(request.param, seed=7)
    yield request.param
    emu.set_schedule(0)


SCENES = [(300, 64, 48, 0, "synthetic", 0.0), (1500, 100, 70, 1, "gflow", 0.25), (800, 33, 17, 2, "synthetic", 1.0)]


@pytest.mark.parametrize("N,W,H,seed,profile,bg", SCENES)
def test_operator_chain_matches_oracle(N, W, H, seed, profile, bg, schedule):
    sc = make_scene(N, W, H, seed=seed, profile=profile, bg=bg)
    Gimg = make_grad_image(3, W, H, seed=seed + 1)
    o = _oracle(sc, Gimg)
    r = emu.operator_chain(sc, Gimg)
    assert torch.equal(r["cov3d"], o["cov3d"]) and torch.equal(r["tiles"], o["tiles"])
    _check(r, o, "operator chain")


@pytest.mark.parametrize("N,W,H,seed,profile,bg", SCENES)
def test_fused_pipeline_matches_oracle(N, W, H, seed, profile, bg, schedule):
    sc = make_scene(N, W, H, seed=seed, profile=profile, bg=bg)
    Gimg = make_grad_image(3, W, H, seed=seed + 1)
    o = _oracle(sc, Gimg)
    r = emu.fused_pipeline(sc, Gimg)
    assert r["rc"] == 0
    _check(r, o, "fused pipeline")


def test_dense_tiles_take_every_sort_path(schedule):
    """Many splats per tile: segments of <=64, <=128, <=256, <=512, <=1024 and >1024 keys (every branch of
    sort_tile_cta), several TMA batches per tile and early termination (opaque splats)."""
    seen = set()
    for N, W, H, seed in [(4000, 48, 32, 5), (2000, 48, 32, 8), (2500, 112, 80, 6), (700, 96, 64, 7)]:
        sc = make_scene(N, W, H, seed=seed, profile="synthetic")
        sc.opacity.clamp_(min=0.6)
        Gimg = make_grad_image(3, W, H, seed=9)
        o = _oracle(sc, Gimg)
        counts = (o["tile_range"][:, 1] - o["tile_range"][:, 0])
        for c in counts.tolist():
            seen.add(next(i for i, lim in enumerate([64, 128, 256, 512, 1024, 1 << 30]) if c <= lim))
        _check(emu.operator_chain(sc, Gimg), o, "dense operator chain")
        _check(emu.fused_pipeline(sc, Gimg), o, "dense fused pipeline")
    assert seen == {0, 1, 2, 3, 4, 5}, seen


@pytest.mark.parametrize("C", [1, 2, 4, 5])
def test_channel_groups(C):
    N, W, H = 400, 48, 40
    sc = make_scene(N, W, H, seed=3, bg=0.5)
    gen = torch.Generator().manual_seed(11)
    feature = torch.rand(N, C, generator=gen)
    Gimg = make_grad_image(C, W, H, seed=4)
    o = _oracle(sc, Gimg, feature)
    _check(emu.operator_chain(sc, Gimg, feature=feature), o, f"C={C} operator chain")
    if C <= 4:
        _check(emu.fused_pipeline(sc, Gimg, feature=feature), o, f"C={C} fused pipeline")


def test_fused_pipeline_capacity_overflow_is_reported():
    sc = make_scene(500, 64, 64, seed=1)
    Gimg = make_grad_image(3, 64, 64)
    full = emu.fused_pipeline(sc, Gimg)
    assert full["rc"] == 0 and full["K"] > 100
    r = emu.fused_pipeline(sc, Gimg, capacity=full["K"] // 2)
    assert r["rc"] == -3 and r["K"] == full["K"]  # GFB_E_CAPACITY, K still delivered for the retry


def test_empty_inputs():
    sc = make_scene(10, 32, 32, seed=0, bg=0.3)
    for k in ("xyz", "scale", "rotate", "opacity", "rgb"):
        setattr(sc, k, getattr(sc, k)[:0].contiguous())
    Gimg = make_grad_image(3, 32, 32)
    r = emu.fused_pipeline(sc, Gimg, capacity=0)
    assert r["rc"] == 0 and r["K"] == 0
    assert torch.allclose(r["image"], torch.full_like(r["image"], 0.3))
    assert int(r["tile_range"].abs().sum()) == 0


@pytest.mark.parametrize("K", [1, 4, 9, 16])
def test_compute_sh_matches_oracle(K):
    """gfb_compute_sh_fwd / bwd (geometry.cu) for degrees 0..3 with a visibility mask and N not a multiple of the CTA."""
    N, Cn = 517, 3
    g = torch.Generator().manual_seed(K)
    shs, dirs = torch.randn(N, Cn, K, generator=g), torch.randn(N, 3, generator=g)
    vis = torch.rand(N, generator=g) > 0.2
    g_out = torch.randn(N, Cn, generator=g)
    L = emu.load()
    for visible in (None, vis):
        v8 = None if visible is None else visible.to(torch.uint8)
        out = emu.f32(N, Cn)
        emu.ok(L.gfb_compute_sh_fwd(emu.p(shs), emu.p(dirs), emu.p(v8), N, Cn, K, emu.p(out), None), "sh fwd")
        ref = C.compute_sh(shs, dirs, visible)
        assert_close(out, ref, 1e-6, "sh forward")
        d_shs, d_dirs = emu.f32(N, Cn, K), emu.f32(N, 3)
        emu.ok(L.gfb_compute_sh_bwd(emu.p(shs), emu.p(dirs), emu.p(v8), N, Cn, K, emu.p(g_out), emu.p(d_shs), emu.p(d_dirs), None),
               "sh bwd")
        r_shs, r_dirs = C.compute_sh_bwd(shs, dirs, visible, g_out)
        assert_close(d_shs, r_shs, 1e-5, "d_shs")
        assert_close(d_dirs, r_dirs, 1e-4, "d_dirs")
    assert L.gfb_compute_sh_fwd(emu.p(shs), emu.p(dirs), None, N, Cn, 5, emu.p(out), None) == -1


def test_randomised_edge_cases():
    """A bounded slice of the fuzz campaign run during development (500 trials, none failing): tiny / odd image sizes
    (down to 1x1), N = 1, splats covering the whole frame, opacities 0 / 1 / 0.003 / 0.999, flat Gaussians, mostly
    culled scenes, 1-4 channels, all three thread schedules -- operator chain and fused pipeline against the oracle."""
    import random

    rnd = random.Random(1234)
    for trial in range(24):
        N = rnd.choice([1, 2, 7, 33, 100, 257, 600])
        W, H = rnd.choice([1, 5, 16, 17, 31, 48, 64, 100]), rnd.choice([1, 3, 16, 20, 33, 47])
        seed, bg = rnd.randrange(10000), rnd.choice([0.0, 0.5, 1.0])
        sc = make_scene(N, W, H, seed=seed, profile=rnd.choice(["synthetic", "gflow"]), bg=bg,
                        outside_frac=rnd.choice([0.0, 0.05, 0.5]))
        mode = rnd.choice(["plain", "huge", "opaque", "transparent", "mixed", "flat"])
        g = torch.Generator().manual_seed(seed)
        if mode == "huge":
            sc.scale = sc.scale * 30
        elif mode == "opaque":
            sc.opacity = torch.full_like(sc.opacity, 0.999)
        elif mode == "transparent":
            sc.opacity = torch.full_like(sc.opacity, 0.003)
        elif mode == "mixed":
            sc.opacity = torch.where(torch.rand(N, 1, generator=g) < 0.5, torch.zeros(N, 1), torch.ones(N, 1))
            sc.scale = sc.scale * torch.where(torch.rand(N, 1, generator=g) < 0.3, 40.0, 1.0)
        elif mode == "flat":
            sc.scale[:, 2] = 1e-9
        Cn = rnd.choice([1, 2, 3, 4])
        feat = torch.rand(N, Cn, generator=g)
        Gimg = make_grad_image(Cn, W, H, seed=seed + 1)
        emu.set_schedule(rnd.choice([0, 1, 2]), seed)
        try:
            o = _oracle(sc, Gimg, feat)
            what = f"trial {trial} N={N} {W}x{H} seed={seed} bg={bg} {mode} C={Cn}"
            _check(emu.operator_chain(sc, Gimg, feature=feat), o, what + " chain")
            _check(emu.fused_pipeline(sc, Gimg, feature=feat), o, what + " fused")
        finally:
            emu.set_schedule(0)


def test_kept_workspaces_clean_themselves():
    """gfb_render_forward_keep / gfb_render_backward_keep: the caller's control block and gradient pack are zero
    before the first call and return to zero by themselves (scatter hands the tile counters back, the scan resets its
    ticket, geometry_bwd clears each pack row after reading it, the blend backward clears d_cam) -- no memset launch
    per step.  Three calls over the same kept blocks, different scenes and a capacity overflow in between, give the
    results of the self-contained entry points."""
    kept = {}
    for i, (N, W, H, seed, cap) in enumerate([(900, 64, 48, 3, None), (900, 64, 48, 4, 50), (900, 64, 48, 5, None)]):
        sc = make_scene(N, W, H, seed=seed, bg=0.2)
        Gimg = make_grad_image(3, W, H, seed=seed)
        a = emu.fused_pipeline(sc, Gimg, capacity=cap, kept=kept)
        b = emu.fused_pipeline(sc, Gimg, capacity=cap)
        assert a["rc"] == b["rc"] and a["K"] == b["K"]
        ctl = kept["control"].view(torch.int32).clone()
        k_word = emu.load().gfb_render_control_k_offset(W, H) // 4
        ctl[k_word] = 0  # the K word keeps its value (it is overwritten, never accumulated)
        n_counts = k_word + 3  # counters + control words; the scanned offsets behind them are rewritten every call
        assert int(ctl[:n_counts].abs().sum()) == 0, f"call {i}: the kept control block must be back at zero"
        if a["rc"] != 0:  # GFB_E_CAPACITY: the caller retries larger; the kept block is already clean for that
            continue
        assert torch.equal(a["image"], b["image"]) and torch.equal(a["ids"], b["ids"])
        for k in a["grads"]:
            assert_close(a["grads"][k], b["grads"][k], 1e-5, f"kept-workspace grad {k}")
        assert int((kept["pack"] != 0).sum()) == 0, f"call {i}: the kept gradient pack must be back at zero"


def test_full_size_config2_fused_pipeline():
    """BASELINE config 2 at full size (60 000 Gaussians, 854x480, K = 197 461) through the shim: ids / tile_range /
    per-Gaussian geometry bit-exact, image 1e-4, gradients 1e-3 against the C oracle (about 20 s)."""
    sc = make_scene(60000, 854, 480, seed=0, profile="synthetic")
    Gimg = make_grad_image(3, 854, 480)
    o = _oracle(sc, Gimg)
    r = emu.fused_pipeline(sc, Gimg, capacity=400000)
    assert r["rc"] == 0 and r["K"] == 197461
    _check(r, o, "cfg2 fused pipeline")


def test_tight_tile_culling_keeps_images_and_gradients():
    """GFB_TIGHT_TILES=1 (the shipped default; fused pipeline + native fit loop only): tests/simt/tight_tiles_check.py in
    a process of its own (the rest of the emulated suite runs with the 3-sigma rule to compare ids bit for bit)."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, os.path.join(here, "simt", "tight_tiles_check.py")], capture_output=True, text=True,
                         env=dict(os.environ, GFB_TIGHT_TILES="1"), cwd=os.path.dirname(here), timeout=900)
    assert res.returncode == 0 and "TIGHT_TILES_OK" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_sanitized_build_finds_no_out_of_bounds_or_misaligned_access():
    """The emulated kernels rebuilt with AddressSanitizer + UBSan (alignment, bounds), aborting on the first finding, and
    run over odd sizes -- the CPU counterpart of compute-sanitizer memcheck: host tensors get red zones (libasan is
    preloaded), so a kernel writing one element past a caller's buffer, or reading a float4 through a pointer that is
    only 8-byte aligned (fine on x86, a fault on the GPU), stops the run."""
    import subprocess

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
    import build_emu

    preload = build_emu.sanitizer_preload()
    if preload is None:
        pytest.skip("libasan / libubsan not available")
    lib = build_emu.build_sanitized()
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, LD_PRELOAD=preload, GFB_EMU_LIB=lib, UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1",
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:detect_stack_use_after_return=0")
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", os.path.join(here, "test_simt_kernels.py"),
                          os.path.join(here, "test_simt_fit.py"), os.path.join(here, "test_simt_densify.py"), "-k",
                          "randomised or channel_groups or every_term or masks or moving_footprint or densif or compute_sh or empty "
                          "or rolled_back or dense_tiles"],
                         capture_output=True, text=True, env=env, cwd=os.path.dirname(here), timeout=1800)
    assert res.returncode == 0 and " passed" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
    assert "runtime error" not in res.stderr and "AddressSanitizer" not in res.stderr


def test_host_pipe_orders_copy_compute_copy():
    """csrc/hostpipe.cu under the shim (copies happen at once, the "graph" is a host callback): every submitted step
    sees its own input in its slot's device block and its own result lands in the host block it named, slots are reused
    round-robin, bad arguments are refused."""
    import ctypes

    import numpy as np

    L = emu.load()
    pipe = ctypes.c_void_p()
    assert L.gfb_hostpipe_create(0, ctypes.byref(pipe)) != 0
    assert L.gfb_hostpipe_create(2, ctypes.byref(pipe)) == 0
    n = 64
    dev_in = [np.zeros(n, np.float32) for _ in range(2)]
    dev_out = [np.zeros(n + 1, np.float32) for _ in range(2)]
    current = {"slot": 0}

    @ctypes.CFUNCTYPE(None)
    def graph():  # stands for the captured render step: reads the slot's input block, writes its output block
        k = current["slot"]
        dev_out[k][:n] = 2.0 * dev_in[k]
        dev_out[k][n] = dev_in[k].sum()

    exec_ptr = ctypes.cast(graph, ctypes.c_void_p)
    outs = []
    for step in range(5):
        k = step % 2
        current["slot"] = k
        host_in = np.full(n, float(step + 1), np.float32)
        host_out = np.zeros(n + 1, np.float32)
        rc = L.gfb_hostpipe_submit(pipe, k, dev_in[k].ctypes.data, host_in.ctypes.data, 4 * n, exec_ptr, None,
                                   host_out.ctypes.data, dev_out[k].ctypes.data, 4 * (n + 1))
        assert rc == 0
        outs.append(host_out)
    assert L.gfb_hostpipe_wait(pipe) == 0
    for step, o in enumerate(outs):
        assert np.all(o[:n] == 2.0 * (step + 1)) and o[n] == n * (step + 1)
    assert L.gfb_hostpipe_submit(pipe, 2, None, None, 0, exec_ptr, None, None, None, 0) != 0  # no such slot
    assert L.gfb_hostpipe_submit(pipe, 0, None, None, 0, None, None, None, None, 0) != 0      # no graph
    assert L.gfb_hostpipe_destroy(pipe) == 0
