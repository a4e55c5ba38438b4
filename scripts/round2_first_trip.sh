#!/bin/bash
# First GPU trip of round 2 (1 GPU, ~2 min): the measurements DESIGN.md section 8 asks for before any kernel work.
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/bulk_red.bin scripts/micro/bulk_red_f64.cu   (here, before gpurun)
mkdir -p gpurun_out
timeout 120 scripts/micro/bulk_red.bin > gpurun_out/r2_bulk_red.txt 2>&1; echo "bulk_red rc=$?"; head -12 gpurun_out/r2_bulk_red.txt
timeout 200 python scripts/gpu_dgemm_flags.py 0 128 256 512 16 > gpurun_out/r2_dgemm_flags.txt 2>&1; echo "flags rc=$?"; cat gpurun_out/r2_dgemm_flags.txt
timeout 120 python scripts/gpu_potrf_bench.py > gpurun_out/r2_potrf.txt 2>&1; echo "potrf rc=$?"; grep "potrf d\|trsm" gpurun_out/r2_potrf.txt
timeout 400 python bench.py > gpurun_out/r2_bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n1.log | cut -c1-400
ELB200_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_el_siblings_gpu.py -m gpu -x -q > gpurun_out/r2_siblings.log 2>&1; echo "siblings rc=$?"; tail -5 gpurun_out/r2_siblings.log
# then, on 4 and 8 GPUs: ELB200_P2P=1 python -m pytest tests/test_multigpu.py -m gpu -x -q   (parity of the peer-memory path on 2x2 / 2x4)
