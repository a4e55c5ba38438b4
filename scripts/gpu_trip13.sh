#!/bin/bash
# trip 13 (1 GPU): full default bench with the new configs[3]/[4] sections, ncu --set full of the tf32 kernel,
# launch list of a reduced bench
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python bench.py > gpurun_out/t13_bench_n1.log 2>&1; echo "bench n1 rc=$? $(( $(date +%s)-S ))s"
tail -1 gpurun_out/t13_bench_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in ('value','dpotrf','zhpdsolve','sgemm_dot','e2e'): print(k, d.get(k))
"
grep -i "error\|Traceback" gpurun_out/t13_bench_n1.log | head -5
S=$(date +%s)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_3xtf32 -s 1 -c 1 -o gpurun_out/t13_prof_tf32 python scripts/gpu_tf32_probe.py perf N N > gpurun_out/t13_ncu_tf32.log 2>&1; echo "ncu tf32 rc=$? $(( $(date +%s)-S ))s"
S=$(date +%s)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t13_launches.csv python bench.py --n 8192 --potrf-n 8192 --hpd-n 4096 --hpd-rhs 256 --sgemm-mn 2048 --sgemm-k 32768 --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/t13_ncu_bench.log 2>&1; echo "ncu list rc=$? $(( $(date +%s)-S ))s"
