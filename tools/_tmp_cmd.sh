timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_native_fit_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
for prof in synthetic gflow; do
python bench.py --steps 50 --warmup 5 --quick --no-cpu-baseline --profile $prof 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$prof value', round(d['value'],1), 'graphed', round(d['graphed']['value'],1), 'inflight', round(d['graphed']['frames_in_flight']['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
python tools/bench_fit.py --native --iters 300 2>/dev/null | tail -1 | cut -c1-200
python tools/bench_fit.py --native --ssim --iters 300 2>/dev/null | tail -1 | cut -c1-200
