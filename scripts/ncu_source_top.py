"""Top stalled SASS instructions of an .ncu-rep source page: python scripts/ncu_source_top.py rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = data[i]
    st = {h[6:]: int(r[idx[h]]) for h in stall_cols if int(r[idx[h]] or 0) > 0}
    print(f"{i:5d} {r[idx['# Samples']]:>6} {r[idx['Source']].strip():<70} {st}")
