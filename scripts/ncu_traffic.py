"""Extract per-launch DRAM traffic + headline metrics of the captured kernel into a small JSON:
python scripts/ncu_traffic.py rep.ncu-rep m n k out.json"""
import csv, io, json, subprocess, sys
rep, m, n, k, outp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
def num(key):
    u, v = d[key]
    v = float(v.replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * mult
res = {"kernel": d["Kernel Name"][1], "m": m, "n": n, "k": k,
       "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
       "duration_s_under_ncu": num("gpu__time_duration.sum"),
       "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][1]),
       "l2_hit_rate_pct": float(d["lts__t_sector_hit_rate.pct"][1]),
       "registers_per_thread": int(float(d["launch__registers_per_thread"][1])),
       "algorithmic_bytes": 8.0 * (m * k + k * n + 2.0 * m * n), "flops": 2.0 * m * n * k}
res["dram_bytes_per_launch"] = res["dram_bytes_read"] + res["dram_bytes_write"]
json.dump(res, open(outp, "w"), indent=1)
print(json.dumps(res))
