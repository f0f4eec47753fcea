#!/bin/bash
# Scaling run on one box with 8 GPUs: bench.py at N = 8, 4, 2, 1 and the frame-sharded Adam loop (config 4).
TAG=${1:-r16}
OUT=gpurun_out; mkdir -p $OUT
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 100 --warmup 10 > $OUT/scale_${TAG}_n$n.json 2> $OUT/scale_${TAG}_n$n.err; echo "N=$n rc=$?"; cat $OUT/scale_${TAG}_n$n.json | cut -c1-400
done
python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > $OUT/scale_${TAG}_n1.json 2> $OUT/scale_${TAG}_n1.err; echo "N=1 rc=$?"; cut -c1-400 $OUT/scale_${TAG}_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/bench_fit.py --iters 300 --frames 48 > $OUT/fit_cfg4_${TAG}_n8.json 2> $OUT/fit_cfg4_${TAG}_n8.err; echo "cfg4 N=8 rc=$?"; cat $OUT/fit_cfg4_${TAG}_n8.json
python tools/bench_fit.py --iters 300 --frames 8 > $OUT/fit_cfg4_${TAG}_n1.json 2> $OUT/fit_cfg4_${TAG}_n1.err; echo "cfg4 N=1 (8 frames) rc=$?"; cat $OUT/fit_cfg4_${TAG}_n1.json
# the same frame-sharded loop with the native iteration (csrc/fit.cu)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/bench_fit.py --iters 300 --frames 48 --native > $OUT/fit_cfg4_native_${TAG}_n8.json 2> $OUT/fit_cfg4_native_${TAG}_n8.err; echo "cfg4 native N=8 rc=$?"; cat $OUT/fit_cfg4_native_${TAG}_n8.json
python tools/bench_fit.py --iters 300 --frames 8 --native > $OUT/fit_cfg4_native_${TAG}_n1.json 2> $OUT/fit_cfg4_native_${TAG}_n1.err; echo "cfg4 native N=1 (8 frames) rc=$?"; cat $OUT/fit_cfg4_native_${TAG}_n1.json
