"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [extra substrings...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
        'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum ', 'sm__throughput.avg.pct',
        'lts__throughput.avg.pct', 'smsp__issue_active.avg.pct', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp', 'smsp__inst_executed.sum ',
        'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64', 'smsp__cycles_active.avg ', 'l1tex__throughput.avg.pct',
        'gpu__compute_memory_throughput', 'sm__sass_inst_executed_op_shared', 'smsp__pcsamp_warps_issue_stalled'] + extra
for vals in rows[2:]:
    print("=" * 60)
    for h, u, v in zip(hdr, units, vals):
        if any(w.strip() in h for w in want) and v not in ("", "0"):
            print(f"{h} [{u}] = {v}")
