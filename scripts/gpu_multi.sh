#!/bin/bash
# Multi-GPU trip: parity worker (optionally with the peer-memory path), then bench A/B.  usage: gpu_multi.sh N [tag]
N=${1:-2}; TAG=${2:-n$N}
mkdir -p gpurun_out
H=$([ "$N" -ge 4 ] && echo 2 || echo 1)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for P in 1 0; do
  S=$(date +%s)
  ELB200_P2P=$P timeout 400 $TR --master-port 2961$P tests/mgpu_worker.py $H > gpurun_out/${TAG}_mgpu_p2p$P.log 2>&1; echo "mgpu p2p=$P rc=$? $(( $(date +%s)-S ))s: $(grep -c 'MGPU OK' gpurun_out/${TAG}_mgpu_p2p$P.log) ok"; grep -i "fail\|error" gpurun_out/${TAG}_mgpu_p2p$P.log | head -5
done
for CFG in ${CFGS:-"1 0 1" "1 0 0" "0 8 0"}; do
  set -- $CFG; P=$1; SMS=$2; MO=$3
  S=$(date +%s)
  ELB200_P2P=$P ELB200_SUMMA_PANEL_SMS=$SMS ELB200_P2P_MEMOPS=$MO timeout 300 $TR --master-port 2963$P bench.py --gpus $N --no-e2e --no-cpu --no-hpdsolve --no-sgemm --no-orient --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_p2p${P}_sms${SMS}_mo$MO.log 2>&1; echo "bench p2p=$P sms=$SMS memops=$MO rc=$? $(( $(date +%s)-S ))s"
  tail -1 gpurun_out/${TAG}_bench_p2p${P}_sms${SMS}_mo$MO.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  dgemm', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['kernel_share_of_step'],3), 'parity', d['parity'].get('NN'), 'dpotrf', round(d['dpotrf']['value']), d['dpotrf'].get('solve_residual'), d['redist'])" || tail -5 gpurun_out/${TAG}_bench_p2p${P}_sms${SMS}_mo$MO.log | cut -c1-300
done
ELB200_P2P=1 ELB200_TRACE=1 timeout 200 $TR --master-port 29650 bench.py --gpus $N --no-e2e --no-cpu --no-hpdsolve --no-sgemm --no-orient --steps 1 --warmup 1 > gpurun_out/${TAG}_trace.log 2>&1; grep "elb200 trace" gpurun_out/${TAG}_trace.log | tail -12
