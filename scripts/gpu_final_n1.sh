#!/bin/bash
# final single-GPU check of a round: the whole GPU test suite, then smoke()
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-S ))s"; tail -6 gpurun_out/final_pytest_gpu.log | cut -c1-250
S=$(date +%s)
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3; echo "smoke $(( $(date +%s)-S ))s"
