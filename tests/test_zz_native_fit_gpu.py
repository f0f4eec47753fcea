"""GPU parity of the native fit iteration (csrc/fit.cu) and of the operator path (gflow_b200/fit.py) against
oracle/fit_ref.py and against what the UNMODIFIED reference trainer recorded (tests/golden/trainer_stages.npz), same
cases as the CPU-emulated run in tests/test_simt_fit.py plus BASELINE config 3's full size.

The checks run in a process of their own (last in the suite) so that a fault in one of them cannot poison the CUDA
context of the other parity tests; a failing case fails the suite."""
import json
import os
import subprocess
import sys

import pytest

import fit_check

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [c[0] for c in fit_check.case_list()] + \
    ["reference_golden_first", "reference_golden_camera", "reference_golden_all",
     "operator_golden_first", "operator_golden_camera", "operator_golden_all", "native_config3_size",
     "native_vs_operator_path", "densification_both_paths", "concurrent_frames_on_streams"]


@pytest.fixture(scope="module")
def results():
    res = subprocess.run([sys.executable, os.path.join(HERE, "gpu_native_fit_runner.py")], capture_output=True, text=True,
                         timeout=1500)
    for line in res.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[len("RESULT "):])
    return {"_crash": (res.stdout + res.stderr)[-3000:]}


@pytest.mark.parametrize("name", CASES)
def test_native_fit_on_gpu(results, name):
    assert "_crash" not in results, results.get("_crash")
    assert results.get(name) == "ok", results.get(name)
