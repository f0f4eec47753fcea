#!/bin/bash
# compute-sanitizer passes over a small render step (memcheck, racecheck, initcheck, synccheck)
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/run_steps.py fused 2 cfg1 > $OUT/sanitizer_${tool}_fused.log 2>&1; echo "fused rc=$?"; tail -3 $OUT/sanitizer_${tool}_fused.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/run_steps.py chain 2 cfg1 > $OUT/sanitizer_${tool}_chain.log 2>&1; echo "chain rc=$?"; tail -3 $OUT/sanitizer_${tool}_chain.log
done
