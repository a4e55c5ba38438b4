#!/bin/bash
# trip 12 (8 GPUs): parity on 2x2 and 2x4 grids, bench at N=8 and N=4, Cholesky phase trace at N=8
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "4-2 or 8-2" > gpurun_out/t12_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$? $(( $(date +%s)-S ))s"
tail -4 gpurun_out/t12_pytest_mgpu.log | cut -c1-600
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
  S=$(date +%s)
  timeout 300 $TR bench.py --gpus $N --no-cpu --steps 2 --warmup 3 > gpurun_out/t12_bench_n$N.log 2>&1; echo "bench n$N rc=$? $(( $(date +%s)-S ))s"
  tail -1 gpurun_out/t12_bench_n$N.log | cut -c1-2600
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
ELB200_TRACE=1 timeout 200 $TR bench.py --gpus 8 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/t12_trace_n8.log 2>&1; echo "trace rc=$?"
grep "elb200 trace" gpurun_out/t12_trace_n8.log | tail -40
