#!/bin/bash
# trip 8 (1 GPU): UMMA descriptor probe, tf32 perf on the K-major path
mkdir -p gpurun_out
timeout 60 scripts/micro/umma_probe.bin > gpurun_out/t8_umma_probe.txt 2>&1; echo "probe rc=$?"
timeout 200 python scripts/gpu_tf32_probe.py perf T N > gpurun_out/t8_tf32_perf_TN.log 2>&1; echo "tf32 perf TN rc=$?"; tail -12 gpurun_out/t8_tf32_perf_TN.log
