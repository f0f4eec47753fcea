"""GPU parity of the native fit iteration (csrc/fit.cu) against oracle/fit_ref.py, same cases as the
CPU-emulated run in tests/test_simt_fit.py.

STATUS: these kernels were written after this round's GPU budget was spent; they pass under the CPU SIMT
emulation but have NOT yet run on hardware.  Until a first green hardware run they are marked
xfail(strict=False) -- a pass shows up as XPASS, a failure as xfailed with the message below -- and run in
a process of their own (last in the suite) so that a fault cannot poison the CUDA context of the
established parity tests."""
import json
import os
import subprocess
import sys

import pytest

import fit_check

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="native fit kernels: first hardware run pending (CPU-emulated parity green)")]
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [c[0] for c in fit_check.case_list()] + ["reference_golden_first", "reference_golden_camera", "reference_golden_all"] + ["native_vs_operator_path", "densification_both_paths",
                                                  "concurrent_frames_on_streams"]


@pytest.fixture(scope="module")
def results():
    res = subprocess.run([sys.executable, os.path.join(HERE, "gpu_native_fit_runner.py")], capture_output=True, text=True,
                         timeout=900)
    for line in res.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[len("RESULT "):])
    return {"_crash": (res.stdout + res.stderr)[-3000:]}


@pytest.mark.parametrize("name", CASES)
def test_native_fit_on_gpu(results, name):
    assert "_crash" not in results, results.get("_crash")
    assert results.get(name) == "ok", results.get(name)
