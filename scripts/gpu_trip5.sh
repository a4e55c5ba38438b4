#!/bin/bash
# trip 5 (2 GPUs): parity suite after the overlap refactoring, N=1 bench, N=2 overlap sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/t5_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t5_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t5_pytest.log
timeout 600 python bench.py > gpurun_out/t5_bench_n1.log 2>&1; echo "bench n1 rc=$?"
tail -1 gpurun_out/t5_bench_n1.log | cut -c1-600
ELB200_OVERLAP=0 timeout 600 python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/t5_bench_n1_noov.log 2>&1; echo "bench n1 noov rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
for cfg in "0 -1" "1 4" "1 -1" "1 16"; do
  set -- $cfg
  ELB200_OVERLAP=$1 ELB200_PANEL_SMS=$2 timeout 600 $TR bench.py --gpus 2 --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/t5_bench_n2_ov$1_sm$2.log 2>&1; echo "bench n2 ov=$1 sms=$2 rc=$?"
  tail -1 gpurun_out/t5_bench_n2_ov$1_sm$2.log | cut -c1-400
done
