#!/bin/bash
# One GPU round trip: parity tests, smoke, bench, ncu launch list and one full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu_${TAG}.txt 2>&1
nproc >> $OUT/gpu_${TAG}.txt
echo "== build" ; timeout 600 python __graft_entry__.py > $OUT/build_${TAG}.log 2>&1; echo "build rc=$?"
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_${TAG}.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_${TAG}.log
echo "== bench" ; timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; cat $OUT/bench_${TAG}.json; tail -5 $OUT/bench_${TAG}.err
echo "== bench gflow profile" ; timeout 600 python bench.py --steps 50 --warmup 5 --profile gflow --no-cpu-baseline > $OUT/bench_gflow_${TAG}.json 2>> $OUT/bench_${TAG}.err; cat $OUT/bench_gflow_${TAG}.json
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_${TAG}.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (blend kernels)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 4 -f -o $OUT/prof_blend_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
