"""Walks fit_check.check_operator_stage_against_reference_golden (a GPU test body) through on the CPU under
tests/simt/fake_cuda_run.py: the msplat operators are answered by the oracle, so this checks the operator path's
host logic (gflow_b200/fit.py FrameFitter.train) against what the unmodified reference trainer recorded."""
import fit_check

for stage in ("first", "camera", "all"):
    fit_check.check_operator_stage_against_reference_golden("cuda:0", stage)
    print("stage", stage, "ok", flush=True)
print("OPERATOR_GOLDEN_OK")
