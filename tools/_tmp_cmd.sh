bash tools/gpu_tests.sh r2final
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cfg5 --no-fit-loop > gpurun_out/bench_cfg5_final.json 2> gpurun_out/bench_cfg5_final.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --profile gflow --no-fit-loop --no-cpu-baseline > gpurun_out/bench_gflow_final.json 2> gpurun_out/bench_gflow_final.err; echo "gflow rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 8 -c 4 -f -o gpurun_out/prof_blend_final python tools/run_steps.py fused 6 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/diag_host.py > gpurun_out/diag_host_final.txt 2>&1
timeout 300 python tools/diag_chain.py 100 > gpurun_out/diag_chain_final.txt 2>&1
