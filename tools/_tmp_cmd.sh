V=gflow_b200/_lib/variants
python tools/ab_blend.py --reps 30 base exact 2>&1 | tail -8
python tools/ab_blend.py --reps 30 --profile gflow base exact 2>&1 | tail -8
for lib in $V/libgfb_base.so $V/libgfb_exact.so; do for prof in synthetic gflow; do
GFLOW_B200_LIB=$lib GFLOW_B200_NO_EXT=1 timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline --profile $prof 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('lib=$lib $prof graphed', round(d['graphed']['value'],1), [round(b,3) for b in d['graphed']['blocks_ms'][:3]])"
done; done
