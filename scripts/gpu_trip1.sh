#!/bin/bash
# first GPU trip of the round: tests, probes, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_probe1.py > gpurun_out/probe1.log 2>&1; echo "probe1 rc=$?"
timeout 300 python scripts/gpu_probe2.py > gpurun_out/probe2.log 2>&1; echo "probe2 rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --n 8192 --potrf-n 16384 --cpu-n 4096 --steps 2 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench_small rc=$?"
tail -1 gpurun_out/bench_small.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench_full rc=$?"
tail -1 gpurun_out/bench_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --n 8192 --potrf-n 8192 --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
