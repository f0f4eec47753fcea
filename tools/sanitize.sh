#!/bin/bash
# compute-sanitizer passes over a small render step (fused pipeline, operator chain) and a few iterations of the native
# fit loop: memcheck, racecheck, synccheck
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/run_steps.py fused 2 cfg1 > $OUT/sanitizer_${tool}_fused.log 2>&1; echo "fused rc=$?"; tail -2 $OUT/sanitizer_${tool}_fused.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/run_steps.py chain 2 cfg1 > $OUT/sanitizer_${tool}_chain.log 2>&1; echo "chain rc=$?"; tail -2 $OUT/sanitizer_${tool}_chain.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/fit_small.py small 3 > $OUT/sanitizer_${tool}_fit.log 2>&1; echo "fit rc=$?"; tail -2 $OUT/sanitizer_${tool}_fit.log
done
