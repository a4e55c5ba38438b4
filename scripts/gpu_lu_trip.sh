#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 420 python -m pytest tests/test_el_lu_gpu.py -x -q -k "lu or getrf" > gpurun_out/lu_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -5 gpurun_out/lu_pytest.log | cut -c1-250

for n in 8192 16384 32768; do timeout 300 python scripts/gpu_lu_bench.py $n 128 2 2>&1 | tail -1 | cut -c1-300; done | tee gpurun_out/lu_perf_gen2.txt
