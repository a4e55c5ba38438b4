#!/bin/bash
# Evidence trip (1 GPU): ncu --set full captures of the kernels DESIGN.md quotes + the launch list of a short bench run.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip script-args...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$re -s $skip -c 1 -o gpurun_out/r2_ev_$name "$@" > gpurun_out/r2_ev_$name.log 2>&1; echo "$name rc=$?"
}
cap dgemm_n1_32768x32768x128 gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 32768 32768 128 N N 3 2
cap dgemm_n2_32768x16384x128 gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 32768 16384 128 N N 3 2
cap dgemm_n4_16384x16384x128 gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 16384 16384 128 N N 3 2
cap dgemm_n8_16384x8192x128 gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 16384 8192 128 N N 3 2
cap dtrrk_L_32768x32768x256_NT gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 32768 32768 256 N T 3 2 trrkL
cap zgemm_16384x8192x128 gemm_f64_ws 1 python scripts/gpu_dgemm_one.py 16384 8192 128 N N 3 2 zgemm
cap potrf_d256 potrf_kernel 2 python scripts/gpu_potrf_bench.py
cap trsm_slab trsm_right_slab 2 python scripts/gpu_potrf_bench.py
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_ev_launches_bench.csv python bench.py --n 8192 --potrf-n 16384 --hpd-n 4096 --hpd-rhs 256 --sgemm-mn 1024 --sgemm-k 16384 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ev_launches_bench.log 2>&1; echo "launch list rc=$?"
