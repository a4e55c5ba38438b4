#!/bin/bash
# Final measurement of a round on N GPUs: the multi-GPU parity worker, then the full default bench line.
N=${1:-8}; TAG=${2:-r2final_n$N}
mkdir -p gpurun_out
H=$([ "$N" -ge 4 ] && echo 2 || echo 1)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S=$(date +%s)
timeout 500 $TR --master-port 29811 tests/mgpu_worker.py $H > gpurun_out/${TAG}_mgpu.log 2>&1; echo "mgpu rc=$? $(( $(date +%s)-S ))s: $(grep -c 'MGPU OK' gpurun_out/${TAG}_mgpu.log) ok"; grep -i "fail" gpurun_out/${TAG}_mgpu.log | head -5
S=$(date +%s)
timeout 900 $TR --master-port 29812 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$? $(( $(date +%s)-S ))s"
tail -1 gpurun_out/${TAG}_bench.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('dgemm', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['kernel_share_of_step'],3), 'parity', d['parity'])
for k in ('dpotrf','zhpdsolve','e2e','dgemm_orientations','other_paths','sgemm_dot'): print(k, json.dumps(d.get(k))[:700])
" || tail -5 gpurun_out/${TAG}_bench.log | cut -c1-300
