import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gpuutil as G
from elemental_b200._lib import lib
rng = np.random.default_rng(7)
def op(X, t): return X if t == "N" else X.T
def ev(x): return x + (x & 1)
m, n, k = 4096, 3072, 64
for flags in (0, 1, 0, 1):
  lib().elb200_dgemm_set_debug_flags(flags); print("FLAGS", flags)
  for ta, tb in (("N", "N"),):
    A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), np.float64)
    B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), np.float64)
    C0 = G.rand(rng, m, n, np.float64)
    P = op(A, ta) @ op(B, tb)
    for (alpha, beta) in ((1.0, 1.0), (1.0, 1.0), (1.0, 1.0)):
        dA = G.DevMat(A, ev(A.shape[0] + 2)); dB = G.DevMat(B, ev(B.shape[0] + 4), offset=2); dC = G.DevMat(C0, m + 5, offset=1)
        G.gemm(ta, tb, alpha, dA, dB, beta, dC, k)
        ref = alpha * P + beta * C0
        got = dC.get()
        bad = np.abs(got - ref) > 1e-9
        print(f"{ta}{tb} alpha={alpha} beta={beta}: bad={int(bad.sum())}", flush=True)
        if bad.sum():
            ii, jj = np.nonzero(bad)
            tiles = sorted(set(zip((ii // 128).tolist(), (jj // 64).tolist())))
            for (tm, tn) in tiles[:3]:
                sub = bad[tm*128:(tm+1)*128, tn*64:(tn+1)*64]
                r, c = np.nonzero(sub)
                print("  tile", tm, tn, "nbad", int(sub.sum()), "rows", sorted(set(r.tolist())), "cols", sorted(set(c.tolist())))
                # is got == P-only / C0-only / stale?
                i0, j0 = tm*128 + r[0], tn*64 + c[0]
                print("   got", got[i0, j0], "ref", ref[i0, j0], "P", P[i0, j0], "C0", C0[i0, j0])
                # per-16-k partial sums to see if error equals one k-stage
                Ao, Bo = op(A, ta), op(B, tb)
                parts = [float(Ao[i0, s:s+16] @ Bo[s:s+16, j0]) for s in range(0, k, 16)]
                # where did the wrong `old` come from?
                for q in range(min(len(r), 400)):
                    if q % 37: continue
                    ii0, jj0 = tm*128 + r[q], tn*64 + c[q]
                    old_used = got[ii0, jj0] - P[ii0, jj0]
                    hit = np.argwhere(np.abs(C0 - old_used) < 1e-12)
                    hitP = np.argwhere(np.abs(P + C0 - old_used) < 1e-12)
                    print("    elem", (int(ii0), int(jj0)), "old_used", old_used, "C0 hits", hit[:3].tolist(), "P+C0 hits", hitP[:3].tolist())
