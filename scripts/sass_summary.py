"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md), from the objects
elemental_b200/_build/*.o (sm_100a):  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MN = ["DMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDG", "LDGSTS", "HMMA", "FFMA", "DFMA", "DMUL", "STL", "LDL"]
print("SASS mnemonic counts per kernel object (cuobjdump -sass, sm_100a).  DMMA = FP64 tensor MMA (mma.sync m8n8k4.f64), UTMALDG = TMA load,")
print("UTCHMMA / LDTM / UTCBAR = tcgen05.mma / tcgen05.ld / tcgen05.commit, REDG = red.global, SYNCS = mbarrier ops, STL/LDL = spills.\n")
for obj in sorted(glob.glob(os.path.join(ROOT, "elemental_b200", "_build", "kernels__*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in MN:
                if op.startswith(k):
                    per[cur][k] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print(f"{os.path.basename(obj)}: {len(per)} kernels; totals " + ", ".join(f"{k} {v}" for k, v in tot.items() if v))
    for name, c in per.items():
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(anonymous namespace\)::|elb200::|<unnamed>::", "", short)[:100]
        if any(c.values()):
            print("    " + short + ": " + ", ".join(f"{k} {v}" for k, v in c.items() if v))
    print()
