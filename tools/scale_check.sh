#!/bin/bash
# N = 1 (quick) and N = max on one multi-GPU box: per-N value / graphed / e2e / sequence, as the driver's scaling run does
NMAX=${1:-8}; TAG=${2:-sc}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-fit-loop --no-proxy > $OUT/scale_${TAG}_n1.json 2> $OUT/scale_${TAG}_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NMAX --steps 20 --warmup 5 > $OUT/scale_${TAG}_n$NMAX.json 2> $OUT/scale_${TAG}_n$NMAX.err; echo "n$NMAX rc=$?"; tail -3 $OUT/scale_${TAG}_n$NMAX.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NMAX --steps 20 --warmup 5 --workload cfg5 --quick --no-cpu-baseline > $OUT/scale_${TAG}_cfg5_n$NMAX.json 2> $OUT/scale_${TAG}_cfg5_n$NMAX.err; echo "cfg5 n$NMAX rc=$?"
python - <<P
import json
a=json.load(open('$OUT/scale_${TAG}_n1.json')); b=json.load(open('$OUT/scale_${TAG}_n$NMAX.json'))
n=$NMAX
for k in ('value',):
    print(k, a[k], b[k], 'eff', b[k]/(n*a[k]))
for k in ('graphed','e2e','operator_chain','gflow_iteration'):
    print(k, a[k]['value'], b[k]['value'], 'eff', b[k]['value']/(n*a[k]['value']))
print('sequence', a['sequence']['value'], b['sequence']['value'], 'speed-up', b['sequence']['value']/a['sequence']['value'], a['sequence']['seconds'], b['sequence']['seconds'])
print('collectives', b['collectives_ms'])
print('blocks n1', a['blocks_ms']); print('blocks nmax', b['blocks_ms'])
print('cores/rank', b['config'].get('host_cores_per_rank'))
try:
    c=json.load(open('$OUT/scale_${TAG}_cfg5_n$NMAX.json')); print('cfg5 at n', n, 'value', c['value'], 'e2e', c['e2e']['value'], 'ms/step', c['ms_per_step'])
except Exception as ex: print('cfg5', ex)
print('inflight', a['graphed'].get('frames_in_flight',{}).get('value'), b['graphed'].get('frames_in_flight',{}).get('value'))
P
