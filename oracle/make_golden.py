"""Generate tests/golden/ref_vectors.npz from the REFERENCE (oracle/_ref/libElRef.so).
Inputs are the seeded hash generator, so only reference OUTPUTS are stored.
Run from the repo root in the container where /root/reference is mounted:
    make -C oracle/refbuild && python oracle/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import elemental_oracle as O  # noqa: E402
from oracle import reference_lib as R  # noqa: E402

out = {}
for dt in (np.float64, np.complex128, np.float32):
    tag = np.dtype(dt).char
    m, n, k, nb = 48, 40, 36, 16
    A, B, C0 = O.fill(0, m, k, 1, dtype=dt), O.fill(0, k, n, 2, dtype=dt), O.fill(0, m, n, 3, dtype=dt)
    for alg in (1, 2, 3, 4):
        out[f"gemm_{tag}_NN_{alg}"] = R.gemm("N", "N", 3.0, A, B, 4.0, C0.copy(order="F"), nb=nb, alg=alg)
for dt in (np.float64, np.complex128):
    tag = np.dtype(dt).char
    n = 60
    A = O.fill(1, n, n, 5, diag=float(n), dtype=dt)
    for uplo in "LU":
        out[f"chol_{tag}_{uplo}"] = R.cholesky(uplo, A.copy(order="F"), nb=16)
        out[f"hpd_{tag}_{uplo}"] = R.hpd_solve(uplo, "N", A, O.fill(0, n, 4, 9, dtype=dt), nb=16)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_vectors.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", R.info()["corename"])
