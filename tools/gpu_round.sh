#!/bin/bash
# One GPU round trip: parity tests, smoke, bench, ncu launch list and full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
nproc >> $OUT/gpu_${TAG}.txt
echo "== build" ; timeout 600 python __graft_entry__.py > $OUT/build_${TAG}.log 2>&1; echo "build rc=$?"
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_${TAG}.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_${TAG}.log
echo "== bench" ; timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; cat $OUT/bench_${TAG}.json; tail -5 $OUT/bench_${TAG}.err
echo "== in-situ kernel times"; timeout 300 python tools/kernel_times.py fused 20 cfg2 synthetic flush > $OUT/kernel_times_fused_${TAG}.txt 2>&1; tail -12 $OUT/kernel_times_fused_${TAG}.txt; timeout 300 python tools/kernel_times.py chain 20 cfg2 synthetic flush > $OUT/kernel_times_chain_${TAG}.txt 2>&1; tail -32 $OUT/kernel_times_chain_${TAG}.txt
if [ -z "$QUICK" ]; then
echo "== bench gflow profile" ; timeout 600 python bench.py --steps 50 --warmup 5 --profile gflow --no-cpu-baseline > $OUT/bench_gflow_${TAG}.json 2>> $OUT/bench_${TAG}.err; cat $OUT/bench_gflow_${TAG}.json
echo "== ncu launch list (fused step)" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_fused_${TAG}.csv python tools/run_steps.py fused 4 > $OUT/ncu_list_${TAG}.log 2>&1; echo "rc=$?"
echo "== ncu launch list (operator chain)" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_chain_${TAG}.csv python tools/run_steps.py chain 4 >> $OUT/ncu_list_${TAG}.log 2>&1; echo "rc=$?"
echo "== ncu full (fused step kernels)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"preprocess|scatter|tile_sort|blend_|geometry_bwd" -s 12 -c 6 -f -o $OUT/prof_fused_${TAG} python tools/run_steps.py fused 4 > $OUT/ncu_full_${TAG}.log 2>&1; echo "rc=$?"
echo "== ncu full (chain binning kernels)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bin_|tile_s|pack_" -s 12 -c 6 -f -o $OUT/prof_chain_${TAG} python tools/run_steps.py chain 4 >> $OUT/ncu_full_${TAG}.log 2>&1; echo "rc=$?"
fi
ls -la $OUT | tail -30
