timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "graphed or host_render" 2>&1 | grep -E "passed|failed|Error|error" | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print('default-flags bench: value', d['value'], 'graphed', d['graphed']['value'], 'e2e', d['e2e']['value'], 'steps', d['steps'], 'launches', d['gpu_launches'])"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>/dev/null ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_ref.json
