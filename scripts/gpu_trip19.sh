#!/bin/bash
# trip 19 (2 GPUs): peer-memory redistribution path (ELB200_P2P=1): parity on 1x2 / 2x1, N=2 bench A/B
mkdir -p gpurun_out
S=$(date +%s)
ELB200_P2P=1 timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/t19_pytest_p2p.log 2>&1; echo "pytest p2p rc=$? $(( $(date +%s)-S ))s"
tail -15 gpurun_out/t19_pytest_p2p.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
for P in 1 0; do
  S=$(date +%s)
  ELB200_P2P=$P timeout 200 $TR bench.py --gpus 2 --no-e2e --no-cpu --no-hpdsolve --no-sgemm --steps 2 --warmup 1 > gpurun_out/t19_bench_n2_p2p$P.log 2>&1; echo "bench n2 p2p=$P rc=$? $(( $(date +%s)-S ))s"
  tail -1 gpurun_out/t19_bench_n2_p2p$P.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['dpotrf']['value'], d['dpotrf'].get('solve_residual'), d['redist'])" || tail -5 gpurun_out/t19_bench_n2_p2p$P.log
done
