#!/bin/bash
# trip 10 (1 GPU): BASE32B probe, tf32 checks all orientations, new tf32 tests, tf32 perf
mkdir -p gpurun_out
timeout 60 scripts/micro/umma_probe.bin > gpurun_out/t10_umma_probe.txt 2>&1; echo "probe rc=$?"
for c in "T N" "N N" "N T" "T T"; do
  timeout 100 python scripts/gpu_tf32_probe.py check $c > gpurun_out/t10_tf32_check_${c// /}.log 2>&1; echo "tf32 check $c rc=$?"
  tail -7 gpurun_out/t10_tf32_check_${c// /}.log | cut -c1-160
done
timeout 300 python -m pytest tests -m gpu -x -q -k "tf32" > gpurun_out/t10_pytest_tf32.log 2>&1; echo "pytest tf32 rc=$?"; tail -5 gpurun_out/t10_pytest_tf32.log
timeout 200 python scripts/gpu_tf32_probe.py perf > gpurun_out/t10_tf32_perf.log 2>&1; echo "tf32 perf rc=$?"; grep 3xTF32 gpurun_out/t10_tf32_perf.log
