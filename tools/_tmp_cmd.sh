for rep in 1 2 3; do for c in 1 3; do
python tools/bench_fit.py --native --frames 48 --iters 300 --concurrent $c 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('concurrent $c', round(d['value'],1), round(d['seconds'],3))"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
done; done
