#!/bin/bash
# Round-2 first contact: baseline of the round-1 kernels on this round's box + A/B of the prepared switches.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
nproc > $OUT/nproc_${TAG}.txt
python -c "import os; print(len(os.sched_getaffinity(0)))" >> $OUT/nproc_${TAG}.txt
echo "== pytest gpu"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rxX > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_${TAG}.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; cut -c1-1200 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
for prof in synthetic gflow; do for k in 0 2 4; do
  GFB_BWD_SPARSE=$k timeout 300 python bench.py --steps 100 --warmup 10 --profile $prof --no-cpu-baseline --no-fit-loop > $OUT/bench_sparse${k}_${prof}_${TAG}.json 2>> $OUT/bench_${TAG}.err
  python -c "import json,sys; d=json.load(open('$OUT/bench_sparse${k}_${prof}_${TAG}.json')); print('$prof sparse=$k', round(d['value'],1), 'it/s  median ms', round(d['ms_per_step_median'],4), ' blend_bwd', round(d['roofline']['kernel_ms']*1e3,1), 'us fwd', round(d['roofline']['blend_fwd']['kernel_ms']*1e3,1))"
done; done
for prof in synthetic gflow; do for t in 0 1; do
  GFB_TIGHT_TILES=$t timeout 300 python bench.py --steps 100 --warmup 10 --profile $prof --no-cpu-baseline --no-fit-loop > $OUT/bench_tight${t}_${prof}_${TAG}.json 2>> $OUT/bench_${TAG}.err
  python -c "import json; d=json.load(open('$OUT/bench_tight${t}_${prof}_${TAG}.json')); print('$prof tight=$t', round(d['value'],1), 'it/s  median ms', round(d['ms_per_step_median'],4), 'K', d['config']['K'])"
done; done
for pdl in 0 1; do
  GFB_FIT_PDL=$pdl timeout 300 python tools/bench_fit.py --iters 300 --native > $OUT/fit_cfg3_native_pdl${pdl}_${TAG}.json 2>> $OUT/fit_cfg3_${TAG}.err; cat $OUT/fit_cfg3_native_pdl${pdl}_${TAG}.json
done
echo "== in-situ kernel times"
timeout 300 python tools/kernel_times.py fused 20 cfg2 synthetic flush > $OUT/kernel_times_fused_${TAG}.txt 2>&1; head -14 $OUT/kernel_times_fused_${TAG}.txt
timeout 300 python tools/kernel_times.py chain 20 cfg2 synthetic flush > $OUT/kernel_times_chain_${TAG}.txt 2>&1; head -24 $OUT/kernel_times_chain_${TAG}.txt
timeout 300 python tools/fit_kernel_times.py 20 > $OUT/fit_kernel_times_${TAG}.txt 2>&1; tail -22 $OUT/fit_kernel_times_${TAG}.txt
