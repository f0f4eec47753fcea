"""The SIMT shim checks kernels, so its own semantics are pinned here with tiny purpose-built kernels
(tests/simt/selfcheck.cu): shuffles, scans, ballots with exited lanes, CTA barriers (with exited threads and with
predicate counts), grid-wide atomics, the poison-until-waited rule of emulated cp.async.bulk, deadlock detection."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(build_emu.build_selfcheck())
    L.gfb_emu_set_schedule.argtypes = [ctypes.c_int, ctypes.c_uint]
    return L


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(params=[0, 1, 2], ids=["ascending", "descending", "shuffled"])
def schedule(request, lib):
    lib.gfb_emu_set_schedule(request.param, 3)
    yield
    lib.gfb_emu_set_schedule(0, 0)


def test_warp_sum_and_scan(lib, schedule):
    x = np.random.default_rng(0).standard_normal(3 * 64).astype(np.float32)
    out = np.zeros_like(x)
    lib.sc_warp_sum(ptr(x), ptr(out), 3)
    ref = np.repeat(x.reshape(-1, 32).sum(axis=1, dtype=np.float64), 32)
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-5)
    assert all(len(set(out[w * 32:(w + 1) * 32])) == 1 for w in range(6)), "every lane of a warp holds the same total"
    v = np.arange(1, 33, dtype=np.int32)
    s = np.zeros_like(v)
    lib.sc_scan_up(ptr(v), ptr(s))
    assert np.array_equal(s, np.cumsum(v))


@pytest.mark.parametrize("keep", [32, 20, 9])
def test_collectives_ignore_exited_lanes(lib, schedule, keep):
    ballots = np.zeros(32, dtype=np.uint32)
    src7 = np.full(32, -1, dtype=np.int32)
    lib.sc_partial_exit(keep, ptr(ballots), ptr(src7))
    even = sum(1 << l for l in range(0, keep, 2))
    expect = even ^ (1 << 30) ^ (1 << 31)  # any(lane == 3) and all(lane < keep) over the live lanes
    assert all(int(b) == expect for b in ballots[:keep]) and all(int(b) == 0 for b in ballots[keep:])
    assert all(int(v) == 70 for v in src7[:keep]) and all(int(v) == -1 for v in src7[keep:])


def test_cta_barrier(lib, schedule):
    flags = (np.arange(256) % 3 == 0).astype(np.int32)
    out = np.zeros(256, dtype=np.int32)
    lib.sc_barrier_count(ptr(flags), ptr(out))
    c = int(flags.sum())
    assert np.array_equal(out, c * 1000 + flags[(np.arange(256) + 97) % 256])
    out = np.zeros(256, dtype=np.int32)
    lib.sc_block_exit(ptr(out))
    assert np.array_equal(out[:64], np.full(64, 42)) and not out[64:].any()


def test_grid_atomics(lib, schedule):
    counter = np.zeros(2, dtype=np.int32)
    fsum = np.zeros(1, dtype=np.float32)
    lib.sc_grid_atomics(ptr(counter), ptr(fsum))
    assert counter[0] == 6 * 96 and counter[1] == 102 and fsum[0] == 0.5 * 6 * 96


def test_bulk_copy_is_poison_until_somebody_waits(lib, schedule):
    src = np.arange(64, dtype=np.float32)
    for wait_first in (0, 1):
        early, late = np.zeros(64, dtype=np.float32), np.zeros(64, dtype=np.float32)
        lib.sc_bulk(ptr(src), ptr(early), ptr(late), wait_first)
        assert np.array_equal(late, src)
        if wait_first:
            assert np.array_equal(early, src)
        else:  # at least the first thread to run read the stage before any wait: it must have seen the poison (NaN)
            assert np.isnan(early).any()


def test_deadlock_is_reported_not_hung():
    code = ("import ctypes,sys,numpy as np; sys.path.insert(0, %r); import build_emu; L = ctypes.CDLL(build_emu.build_selfcheck()); "
            "o = np.zeros(4, dtype=np.int32); L.sc_deadlock(o.ctypes.data_as(ctypes.c_void_p))") % os.path.join(
        os.path.dirname(__file__), "simt")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and "deadlock" in res.stderr
