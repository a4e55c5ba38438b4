#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "tma or agree" > gpurun_out/pytest_tma.log 2>&1; echo "pytest tma rc=$?"
tail -15 gpurun_out/pytest_tma.log
timeout 300 python scripts/gpu_dgemm_bench.py > gpurun_out/dgemm_bench.log 2>&1; echo "bench rc=$?"
cat gpurun_out/dgemm_bench.log
