#!/bin/bash
# trip 6 (1 GPU): tcgen05 3xTF32 kernel bring-up (each step under its own timeout), Cholesky phase trace
mkdir -p gpurun_out
for c in "T N" "N N" "N T" "T T"; do
  timeout 120 python scripts/gpu_tf32_probe.py check $c > gpurun_out/t6_tf32_check_${c// /}.log 2>&1; echo "tf32 check $c rc=$?"
  tail -3 gpurun_out/t6_tf32_check_${c// /}.log
done
timeout 300 python scripts/gpu_tf32_probe.py perf T N > gpurun_out/t6_tf32_perf_TN.log 2>&1; echo "tf32 perf TN rc=$?"; cat gpurun_out/t6_tf32_perf_TN.log | tail -12
timeout 300 python scripts/gpu_tf32_probe.py perf N N > gpurun_out/t6_tf32_perf_NN.log 2>&1; echo "tf32 perf NN rc=$?"; cat gpurun_out/t6_tf32_perf_NN.log | tail -12
ELB200_TRACE=1 timeout 600 python bench.py --n 8192 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/t6_trace.log 2>&1; echo "trace rc=$?"
grep "elb200 trace" gpurun_out/t6_trace.log | tail -24
