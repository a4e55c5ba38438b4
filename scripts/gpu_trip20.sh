#!/bin/bash
# trip 20 (8 GPUs): N=8 bench A/B of the peer-memory redistribution path, after the potrf / trsm kernel rewrite
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618"
for P in 1 0; do
  S=$(date +%s)
  ELB200_P2P=$P timeout 150 $TR bench.py --gpus 8 --no-e2e --no-cpu --no-hpdsolve --no-sgemm --steps 2 --warmup 1 > gpurun_out/t20_bench_n8_p2p$P.log 2>&1; echo "bench n8 p2p=$P rc=$? $(( $(date +%s)-S ))s"
  tail -1 gpurun_out/t20_bench_n8_p2p$P.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['dpotrf']['value'], d['dpotrf'].get('solve_residual'), d['redist'])" || tail -5 gpurun_out/t20_bench_n8_p2p$P.log | cut -c1-300
done
