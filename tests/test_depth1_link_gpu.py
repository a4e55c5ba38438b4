"""INTEGRATION.md depth 1, executed: the REFERENCE's own host code (El::Gemm SUMMA loops, El::Cholesky
LowerVariant3Blocked + LocalTrrk recursion + unblocked diagonal block, El::HPDSolve, El::Trsm, El::LU) running unmodified
against libelb200.so's Fortran BLAS symbols.

oracle/_ref/libElRefDev.so is the same set of reference objects as libElRef.so, except that the two translation units
binding BLAS/LAPACK leave ?gemm_/?trsm_/?syrk_/?herk_/?syr_/?her_ unresolved, so they bind to libelb200.so, and
new[] inside that library returns device-mapped memory (what the Memory<G> allocator patch of INTEGRATION.md does).
(?scal_/?axpy_/?lacpy_ stay on OpenBLAS in this build: the reference also calls them on std::vector staging
buffers, which are pageable host memory.)  ELB200_BLAS_SYNC=1 makes the Fortran entry points behave like a
host BLAS (synchronous), since the reference reads its buffers from the host between calls without fences.

Checked: results agree with the plain CPU build (libElRef.so, OpenBLAS) to the Gemm / Cholesky tolerances, and the
GPU kernels really ran (elb200_launch_count)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, {root!r})
os.environ["ELB200_BLAS_SYNC"] = "1"
import torch
torch.cuda.init(); torch.cuda.set_device(0); torch.zeros(1, device="cuda")
from oracle import elemental_oracle as O
from oracle import reference_lib as R
from elemental_b200._lib import lib
L = lib(); L.elb200_launch_count.restype = C.c_ulonglong
def run_all():
    out = {{}}
    m, n, k, nb = 300, 260, 280, 64
    for dt, tag in ((np.float64, "d"), (np.complex128, "z")):
        A, B, C0 = O.fill(0, m, k, 1, dtype=dt), O.fill(0, k, n, 2, dtype=dt), O.fill(0, m, n, 3, dtype=dt)
        for alg in (1, 3):
            out[f"gemm_{{tag}}_{{alg}}"] = R.gemm("N", "N", 3.0, A, B, 4.0, C0.copy(order="F"), nb=nb, alg=alg)
        At = O.fill(0, k, m, 4, dtype=dt)
        out[f"gemm_{{tag}}_CN"] = R.gemm("C", "N", -1.0, At, B, 1.0, C0.copy(order="F"), nb=nb, alg=3)
        H = O.fill(1, 320, 320, 5, diag=320.0, dtype=dt)
        for uplo in "LU":
            out[f"chol_{{tag}}_{{uplo}}"] = R.cholesky(uplo, H.copy(order="F"), nb=nb)
        Bh = O.fill(0, 320, 40, 6, dtype=dt)
        out[f"hpd_{{tag}}"] = R.hpd_solve("L", "N", H, Bh.copy(order="F"), nb=nb)
        T = np.asfortranarray((O.fill(0, 200, 200, 7, dtype=dt) / 200 + 2 * np.eye(200)).astype(dt))
        out[f"trsm_{{tag}}"] = R.trsm("L", "L", "N", "N", 2.0, T, O.fill(0, 200, 50, 8, dtype=dt).copy(order="F"), nb=nb)
        # El::LU without pivoting (LU.cpp:47-99): its LocalTrsm / LocalGemm leaves go to the GPU, lu::Unb stays on the host
        D = np.asfortranarray((O.fill(0, 280, 280, 9, dtype=dt) + 280 * np.eye(280)).astype(dt))
        out[f"lu_{{tag}}"] = R.lu(D.copy(order="F"), nb=nb)
    return out
cpu = run_all()
R.use_device_build(True)
before = int(L.elb200_launch_count(0))
gpu = run_all()
launches = int(L.elb200_launch_count(0)) - before
allocs = int(R.lib().elref_device_allocs())
worst = 0.0
for key in cpu:
    num = np.linalg.norm(gpu[key] - cpu[key]); den = np.linalg.norm(cpu[key])
    worst = max(worst, num / den)
    print(key, num / den)
print(json.dumps({{"launches": launches, "device_allocs": allocs, "worst_rel_diff": worst}}))
"""


def test_reference_host_code_runs_on_the_gpu_through_the_fortran_blas_symbols():
    dev = os.path.join(ROOT, "oracle", "_ref", "libElRefDev.so")
    if not (os.path.exists(dev) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libElRef.so"))):
        pytest.skip("oracle/_ref/libElRefDev.so not built (make -C oracle/refbuild dev)")
    # a fresh interpreter: the two builds of the reference must not share a process with other tests' state
    p = subprocess.run([sys.executable, "-c", WORKER.format(root=ROOT)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    import json
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["launches"] > 100, d          # the GPU kernels ran (hundreds of dgemm_ / dtrsm_ / dsyr_ / zher_ calls)
    assert d["device_allocs"] > 10, d      # the reference's buffers came from the device-addressable allocator
    # both builds run the same algorithm at the same blocksize; only the leaf arithmetic differs
    assert d["worst_rel_diff"] <= 1e-11, p.stdout[-3000:]
