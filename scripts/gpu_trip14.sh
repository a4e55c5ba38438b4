#!/bin/bash
# trip 14 (1 GPU): DMMA potrf kernel (parity + time), trsm time, DGEMM stagger flag, Dot block edge
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "potrf or cholesky or hpdsolve or dot_blocksize" > gpurun_out/t14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t14_pytest.log
timeout 120 python scripts/gpu_potrf_bench.py > gpurun_out/t14_potrf.log 2>&1; echo "potrf bench rc=$?"; cat gpurun_out/t14_potrf.log | tail -12
timeout 200 python scripts/gpu_dgemm_flags.py 0 1 > gpurun_out/t14_flags.log 2>&1; echo "flags rc=$?"; cat gpurun_out/t14_flags.log | tail -6
timeout 300 python bench.py --no-e2e --no-cpu --no-hpdsolve --steps 1 --warmup 3 > gpurun_out/t14_bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/t14_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','dpotrf','sgemm_dot'): print(k, d.get(k))
"
