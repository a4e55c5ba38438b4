#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench_full rc=$?"
tail -1 gpurun_out/bench_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 2 -c 1 -o gpurun_out/prof_dgemm_update_32768 python scripts/gpu_dgemm_one.py 32768 32768 128 N N 3 4 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01.csv python bench.py --n 8192 --potrf-n 8192 --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
