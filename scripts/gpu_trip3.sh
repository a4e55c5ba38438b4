#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 2 -c 1 -o gpurun_out/prof_dgemm_tma_k128 python scripts/gpu_dgemm_one.py 16384 8192 128 N N 3 4 > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 1 -c 1 -o gpurun_out/prof_dgemm_tma_k8192 python scripts/gpu_dgemm_one.py 8192 8192 8192 N N 3 3 > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
