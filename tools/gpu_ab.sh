#!/bin/bash
# A/B of the blend-kernel variants (tools/build_variants.py) in one call
OUT=gpurun_out; TAG=${1:-ab}; mkdir -p $OUT
for wl in "cfg2 synthetic" "cfg2 gflow" "cfg5 synthetic"; do
  set -- $wl
  timeout 600 python tools/ab_blend.py --workload $1 --profile $2 2>&1 | tee -a $OUT/ab_${TAG}.txt
done
timeout 300 python tools/ab_blend.py --workload cfg2 --profile gflow --channels 4 2>&1 | tee -a $OUT/ab_${TAG}.txt
