#!/bin/bash
# trip 7 (1 GPU): parity suite of the committed state, default bench, first tf32 check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/t7_smi.txt 2>&1
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/t7_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"
tail -14 gpurun_out/t7_pytest.log
S=$(date +%s)
timeout 500 python bench.py > gpurun_out/t7_bench_n1.log 2>&1; echo "bench n1 rc=$? $(( $(date +%s)-S ))s"
tail -1 gpurun_out/t7_bench_n1.log | cut -c1-3000
timeout 100 python scripts/gpu_tf32_probe.py check T N > gpurun_out/t7_tf32_check_TN.log 2>&1; echo "tf32 check TN rc=$?"
tail -8 gpurun_out/t7_tf32_check_TN.log
timeout 100 python scripts/gpu_tf32_probe.py check N N > gpurun_out/t7_tf32_check_NN.log 2>&1; echo "tf32 check NN rc=$?"
tail -8 gpurun_out/t7_tf32_check_NN.log
