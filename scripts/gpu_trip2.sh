#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "tma or agree" > gpurun_out/pytest_tma.log 2>&1; echo "pytest tma rc=$?"
tail -3 gpurun_out/pytest_tma.log
timeout 200 python scripts/gpu_debug_nt.py 2>&1 | grep -E "FLAGS|alpha"
timeout 300 python scripts/gpu_dgemm_flags.py $FLAGS > gpurun_out/dgemm_flags.log 2>&1; cat gpurun_out/dgemm_flags.log
