#!/bin/bash
# multi-GPU parity worker on N GPUs (grid height H); extra environment in $3
N=${1:-4}; H=${2:-2}; TAG=${3:-default}
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 tests/mgpu_worker.py $H > gpurun_out/mgpu_n${N}_h${H}_${TAG}.log 2>&1
echo "mgpu rc=$? $(( $(date +%s)-S ))s"; grep -i "MGPU\|fail" gpurun_out/mgpu_n${N}_h${H}_${TAG}.log | head -20 | cut -c1-400
