for c in 1 2 3 4; do
python tools/bench_fit.py --native --frames 12 --iters 300 --concurrent $c 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('concurrent $c', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('iters_per_sec','seconds','value','frame_iterations_per_s','concurrent_frames')} , list(d.keys())[:12])"
done
for c in 1 2 4; do
python tools/bench_fit.py --native --ssim --frames 8 --iters 300 --concurrent $c 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ssim concurrent $c', d.get('seconds'))"
done
