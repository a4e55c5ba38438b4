#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 420 python -m pytest tests/test_el_lu_gpu.py -x -q -k "cholesky" > gpurun_out/lu_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -25 gpurun_out/lu_pytest.log | cut -c1-250
