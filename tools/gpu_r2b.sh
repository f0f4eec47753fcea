#!/bin/bash
# Round-2 kernel iteration: parity + bench + in-situ kernel times (+ optional ncu of the blend kernels).
TAG=${1:-r2b}
NCU=${2:-}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_${TAG}.log
for prof in synthetic gflow; do
  timeout 300 python bench.py --steps 100 --warmup 10 --profile $prof --no-cpu-baseline --no-fit-loop > $OUT/bench_${prof}_${TAG}.json 2>> $OUT/bench_${TAG}.err
  python -c "import json,sys; d=json.load(open('$OUT/bench_${prof}_${TAG}.json')); print('$prof', round(d['value'],1), 'it/s  median ms', round(d['ms_per_step_median'],4), ' chain', round(d['operator_chain']['value'],1), ' e2e', round(d['e2e']['value'],1), ' blend_bwd', round(d['roofline']['kernel_ms']*1e3,1), 'us fwd', round(d['roofline']['blend_fwd']['kernel_ms']*1e3,1))"
done
timeout 300 python bench.py --steps 50 --warmup 10 --workload cfg5 --no-cpu-baseline --no-fit-loop > $OUT/bench_cfg5_${TAG}.json 2>> $OUT/bench_${TAG}.err
python -c "import json,sys; d=json.load(open('$OUT/bench_cfg5_${TAG}.json')); print('cfg5', round(d['value'],1), 'it/s  median ms', round(d['ms_per_step_median'],4), ' blend_bwd', round(d['roofline']['kernel_ms']*1e3,1), 'us fwd', round(d['roofline']['blend_fwd']['kernel_ms']*1e3,1), 'frac', round(d['roofline']['frac'],4))"
GFB_FIT_PDL=1 timeout 300 python tools/bench_fit.py --iters 300 --native > $OUT/fit_cfg3_native_${TAG}.json 2>> $OUT/fit_cfg3_${TAG}.err; cut -c1-200 $OUT/fit_cfg3_native_${TAG}.json
echo "== in-situ kernel times"
timeout 300 python tools/kernel_times.py fused 20 cfg2 synthetic > $OUT/kernel_times_fused_${TAG}.txt 2>&1; grep "us/step" $OUT/kernel_times_fused_${TAG}.txt
if [ -n "$NCU" ]; then
  echo "== ncu full (blend kernels)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 8 -c 4 -f -o $OUT/prof_blend_${TAG} python tools/run_steps.py fused 6 > $OUT/ncu_full_${TAG}.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full_${TAG}.log
fi
