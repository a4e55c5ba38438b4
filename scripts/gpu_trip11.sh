#!/bin/bash
# trip 11 (2 GPUs): multi-GPU parity on 1x2 and 2x1 grids, N=2 bench with and without overlap
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/t11_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$? $(( $(date +%s)-S ))s"
tail -4 gpurun_out/t11_pytest_mgpu.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
S=$(date +%s)
timeout 400 $TR bench.py --gpus 2 --no-cpu > gpurun_out/t11_bench_n2.log 2>&1; echo "bench n2 rc=$? $(( $(date +%s)-S ))s"
tail -1 gpurun_out/t11_bench_n2.log | cut -c1-2500
S=$(date +%s)
ELB200_OVERLAP=0 timeout 300 $TR bench.py --gpus 2 --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/t11_bench_n2_noov.log 2>&1; echo "bench n2 noov rc=$? $(( $(date +%s)-S ))s"
tail -1 gpurun_out/t11_bench_n2_noov.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['dpotrf'], d['redist'])"
