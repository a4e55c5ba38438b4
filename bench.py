#!/usr/bin/env python
"""bench.py -- distributed DGEMM (and DPOTRF) of the Elemental hot path on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     (the reference's own CPU path, rank 0 only)

Workload (BASELINE.json configs[1]): El::Gemm NN double m=n=k=32768, GEMM_SUMMA_C,
Blocksize() 128, DistMatrix<double,MC,MR> on a 1x1 / 1x2 / 2x2 / 2x4 Grid for N = 1/2/4/8.
A "step" is one El::Gemm call (2mnk = 70.4 TFlop); strong scaling (the matrix is fixed).
Flop conventions are the reference's (tests/blas_like/Gemm.cpp:97, tests/lapack_like/
Cholesky.cpp:171).  One JSON line is printed by rank 0.

Nothing in the timed GPU path touches oracle/; the oracle build of the reference
(oracle/_ref/libElRef.so) is used for the `cpu_baseline` leg and for `--impl reference` only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_HEIGHT = {1: 1, 2: 1, 4: 2, 8: 2}
METRIC = "FP64 GFLOP/s (device-timed, max over ranks) for DGEMM/DPOTRF at 1/2/4/8 B200, % peak"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=32768,
                    help="m=n=k of the DGEMM workload (--size under torchrun, whose own parser claims --n)")
    ap.add_argument("--nb", type=int, default=128)
    ap.add_argument("--potrf-n", type=int, default=65536)
    ap.add_argument("--potrf-nb", type=int, default=256)
    ap.add_argument("--no-potrf", action="store_true")
    ap.add_argument("--no-hpdsolve", action="store_true", help="skip configs[3] (ZHPDSolve n=32768, 1024 rhs)")
    ap.add_argument("--no-sgemm", action="store_true", help="skip configs[4] (SGEMM SUMMA_Dot k=262144, FFMA vs 3xTF32)")
    ap.add_argument("--hpd-n", type=int, default=32768)
    ap.add_argument("--hpd-rhs", type=int, default=1024)
    ap.add_argument("--sgemm-mn", type=int, default=8192)
    ap.add_argument("--sgemm-k", type=int, default=262144)
    ap.add_argument("--no-orient", action="store_true", help="skip the NT / TN orientations of configs[1]")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-k", type=int, default=1024,
                    help="summation indices of the bounded CPU sample (the reference runs C += A(:, :cpu_k) B(:cpu_k, :) at the full m = n)")
    ap.add_argument("--e2e-serial", action="store_true", help="also time the unpipelined host path (3 copies in, Gemm, 1 copy out)")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# clocks sampled during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop_flag, self.index = [], False, index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------
# the reference's own CPU path (oracle/_ref), used for cpu_baseline and --impl reference
# ---------------------------------------------------------------------------
def cpu_reference_gemm(n, nb, steps, warmup, ksample):
    """The reference's own El::Gemm (SUMMA_C, nb) on the box's host cores, on a bounded sample of the SAME workload:
    the m = n = `n` update C += A(:, 0:ksample) B(0:ksample, :) -- the first ksample / nb of the n / nb rank-nb panel
    steps of the full product, same C, same panel shape, alpha = beta = 1 as on the GPU arm."""
    from oracle import elemental_oracle as O
    from oracle import reference_lib as R
    cores = os.cpu_count() or 1
    if R.available():
        R.set_threads(cores)
        kind, info = "reference", R.info()
        run = lambda A, B, Cm: R.gemm("N", "N", 1.0, A, B, 1.0, Cm, nb=nb, alg=3)
        threads = R.info()["threads"]
        blas = f"OpenBLAS core {info['corename']}"
    else:
        kind, threads, blas = "port", 1, "numpy"
        run = lambda A, B, Cm: O.gemm("N", "N", 1.0, A, B, 1.0, Cm, nb=nb, alg=3)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.uniform(-1, 1, (n, ksample)))
    B = np.asfortranarray(rng.uniform(-1, 1, (ksample, n)))
    Cm = np.empty((n, n), order="F")
    Cm[:] = 0.5
    for _ in range(warmup):
        run(A, B, Cm)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(A, B, Cm)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": 2.0 * n * n * ksample / dt / 1e9, "unit": "GFLOP/s", "cores": int(threads), "kind": kind,
            "sample": f"El::Gemm NN double GEMM_SUMMA_C nb={nb} on a 1x1 Grid, m=n={n} with the first {ksample} of the "
                      f"{n} summation indices ({ksample // nb} of {n // nb} rank-{nb} panel steps of the same update), "
                      f"alpha=beta=1, {blas}, {warmup} warm-up + {steps} timed call(s), {dt:.2f} s each"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(0, min(args.warmup, 3))
    base, dt = cpu_reference_gemm(args.n, args.nb, steps, warmup, args.cpu_k)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"El::Gemm NN double m=n=k={args.n} GEMM_SUMMA_C nb={args.nb} on DistMatrix<double,MC,MR>",
                       "reference_sample": base["sample"], "grid": "1x1 (CPU)", "alpha": 1.0, "beta": 1.0},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE {world}; using {world}", file=sys.stderr)
    N = world

    from elemental_b200 import api as El
    from elemental_b200._lib import lib
    L = lib()
    L.elb200_launch_count.restype = C.c_ulonglong

    grid = El.Grid(GRID_HEIGHT.get(N, 0))
    r, c = grid.Height(), grid.Width()
    n, nb = args.n, args.nb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA-event time of `steps` calls, barrier + synchronize on both sides, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- FP64 ceilings measured in this run (no FP64 figure in MEASURED_PEAKS.json) ----
    peak = C.c_double()
    L.elb200_dmma_peak(20000, C.byref(peak), None)
    dmma_peak_tf = peak.value / 1e12
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cublas_tf = 2 * 8192 ** 3 / best / 1e9
    del a, b

    # ---- DGEMM workload: inputs resident in HBM (grid-independent hash fill) ----
    El.SetBlocksize(nb)
    A = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 1)
    B = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 2)
    Cm = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 3)
    step = lambda: El.Gemm(El.NORMAL, El.NORMAL, 1.0, A, B, 1.0, Cm, El.GEMM_SUMMA_C)

    timed(step, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.elb200_launch_count(1)
    L.elb200_gemm_profile(1)
    El.RedistStats(reset=True)
    ms = timed(step, args.steps)
    launches = int(L.elb200_launch_count(0))
    kms, kcount, kflops = C.c_double(), C.c_longlong(), C.c_double()
    L.elb200_gemm_profile_read(C.byref(kms), C.byref(kcount), C.byref(kflops))
    L.elb200_gemm_profile(0)
    clocks = sampler.stop() if rank == 0 else None
    stats = El.RedistStats()
    # DRAM traffic of one rank-nb update launch, from the committed ncu --set full capture of the same
    # launch shape (profiles/r01_dgemm_update_traffic.json, written by scripts/ncu_traffic.py)
    traffic = None
    try:
        for fn_ in ("r02_dgemm_update_traffic.json", "r01_dgemm_update_traffic.json"):
            path_ = os.path.join(ROOT, "profiles", fn_)
            if not os.path.exists(path_):
                continue
            for prof in (lambda j: j if isinstance(j, list) else [j])(json.load(open(path_))):
                if prof.get("m") == (n + r - 1) // r and prof.get("n") == (n + c - 1) // c and prof.get("k") == nb:
                    traffic = prof["dram_bytes_per_launch"]
                    break
            if traffic is not None:
                break
    except Exception:
        pass
    flops_step = 2.0 * n ** 3
    value = flops_step * args.steps / (ms * 1e-3) / 1e9
    kernel_tf = kflops.value / (kms.value * 1e-3) / 1e12 if kms.value > 0 else 0.0

    # ---- parity of the timed product, at every N, outside the timed region: the reference's own check
    #      (tests/blas_like/Gemm.cpp:13-42) ||(alpha op(A) op(B) + beta C0) X - C X||_F / ||C X||_F with 16 hash-filled
    #      right-hand sides.  C holds C0 + T op(A) op(B) after T calls with alpha = beta = 1. ----
    def gemm_parity(oa, ob, A, B, Cfinal, T, seedC=3):
        nr = 16
        X = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, nr).HashFill(0, 9)
        Z = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, nr)
        Y = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, nr)
        El.Gemm(ob, El.NORMAL, 1.0, B, X, 0.0, Z)            # Z = op(B) X
        El.Gemm(El.NORMAL, El.NORMAL, 1.0, Cfinal, X, 0.0, Y)  # Y = C X
        den = El.FrobeniusNorm(Y)
        El.Gemm(oa, El.NORMAL, -float(T), A, Z, 1.0, Y)        # Y -= T op(A) Z
        C0 = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, seedC)
        El.Gemm(El.NORMAL, El.NORMAL, -1.0, C0, X, 1.0, Y)     # Y -= C0 X
        res = El.FrobeniusNorm(Y) / den if den > 0 else float("nan")
        del C0, X, Z, Y
        torch.cuda.empty_cache()
        return float(res)

    parity = {"def": "||(alpha op(A) op(B) + beta C0) X - C X||_F / ||C X||_F, 16 hash-filled right-hand sides "
                     "(the reference's check, tests/blas_like/Gemm.cpp:13-42), C after all warm-up + timed calls",
              "tolerance": 100 * n * 2.0 ** -52}
    try:
        parity["NN"] = gemm_parity(El.NORMAL, El.NORMAL, A, B, Cm, args.warmup + args.steps)
        parity["ok"] = bool(parity["NN"] <= parity["tolerance"])
    except Exception as ex:
        parity["error"] = repr(ex)[:200]

    # ---- DPOTRF (BASELINE.json configs[2]) ----
    potrf = None
    if not args.no_potrf:
        del A, B, Cm
        torch.cuda.empty_cache()
        pn, pnb = args.potrf_n, args.potrf_nb
        El.SetBlocksize(pnb)
        H = El.DistMatrix(np.float64, El.MC, El.MR, grid, pn, pn)
        def potrf_step():
            H.HashFill(1, 5, float(pn))   # refill (A = S + n I) -- not counted: timed separately below
        def factor():
            El.Cholesky(El.LOWER, H)
        times = []
        for it in range(1 + max(1, min(args.steps, 2))):
            potrf_step()
            t = timed(factor, 1)
            if it > 0:
                times.append(t)
        pms = sum(times) / len(times)
        potrf = {"workload": f"El::Cholesky LOWER double HPD n={pn} nb={pnb}", "ms": pms,
                 "value": (pn ** 3 / 3.0) / (pms * 1e-3) / 1e9, "unit": "GFLOP/s",
                 "frac_of_dmma_peak": (pn ** 3 / 3.0) / (pms * 1e-3) / 1e12 / (dmma_peak_tf * N)}
        # solve check on the factor just computed (the reference's own acceptance test,
        # tests/lapack_like/Cholesky.cpp:47-82): X = A \ Y, then ||A X - Y||_F / (n eps ||A||_F ||X||_F)
        try:
            nrhs = 16
            Y = El.DistMatrix(np.float64, El.MC, El.MR, grid, pn, nrhs).HashFill(0, 7)
            X = El.DistMatrix(np.float64, El.MC, El.MR, grid, pn, nrhs).HashFill(0, 7)
            El.CholeskySolveAfter(El.LOWER, El.NORMAL, H, X)
            H.HashFill(1, 5, float(pn))          # the original A again
            El.Gemm(El.NORMAL, El.NORMAL, -1.0, H, X, 1.0, Y)
            potrf["solve_residual"] = El.FrobeniusNorm(Y) / (pn * 2.0 ** -52 * El.FrobeniusNorm(H) * El.FrobeniusNorm(X))
            potrf["solve_residual_def"] = "||A X - Y||_F / (n eps ||A||_F ||X||_F), X from cholesky::SolveAfter, 16 rhs"
            del Y, X
        except Exception as ex:  # the check must never cost the timing line
            potrf["solve_residual_error"] = repr(ex)[:200]
        del H
        torch.cuda.empty_cache()
        El.SetBlocksize(nb)

    # ---- end-to-end: HOST buffers in, HOST result out, through the public API (El.GemmHost = ElGemmDistHost_d):
    #      the [MC,MR] local matrices live in pinned host memory, as a reference DistMatrix's buffers do ----
    e2e = None
    if not args.no_e2e:
        A = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 1)
        B = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 2)
        Cm = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 3)
        lh, lw = A.LocalHeight(), A.LocalWidth()
        pinned = [torch.empty((lw, max(lh, 1)), dtype=torch.float64, pin_memory=True) for _ in range(3)]
        for M, t in zip((A, B, Cm), pinned):
            L.ElDistMatrixLocalToHost_d(M._h, C.c_void_p(t.data_ptr()), max(lh, 1))
        serial = None
        if args.e2e_serial:
            out_pinned = torch.empty((lw, max(lh, 1)), dtype=torch.float64, pin_memory=True)

            def serial_step():
                for M, t in zip((A, B, Cm), pinned):
                    L.ElDistMatrixLocalFromHost_d(M._h, C.c_void_p(t.data_ptr()), max(lh, 1))
                El.Gemm(El.NORMAL, El.NORMAL, 1.0, A, B, 1.0, Cm, El.GEMM_SUMMA_C)
                L.ElDistMatrixLocalToHost_d(Cm._h, C.c_void_p(out_pinned.data_ptr()), max(lh, 1))
            timed(serial_step, 1)
            sms = timed(serial_step, 1)
            serial = {"value": flops_step / (sms * 1e-3) / 1e9, "ms_per_step": sms,
                      "what": "3 x LocalFromHost, El::Gemm, LocalToHost back to back (round 1's e2e)"}
            del out_pinned
        del A, B, Cm
        torch.cuda.empty_cache()

        def e2e_step():
            El.GemmHost(El.NORMAL, El.NORMAL, 1.0, grid, n, n, n, pinned[0], pinned[1], 1.0, pinned[2], El.GEMM_SUMMA_C)

        e2e_steps = max(1, min(args.steps, 2))
        timed(e2e_step, 1)
        ems = timed(e2e_step, e2e_steps)
        bytes_local = lh * lw * 8
        e2e = {"value": flops_step * e2e_steps / (ems * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(3 * bytes_local * N), "d2h_bytes_per_step": int(bytes_local * N),
               "ms_per_step": ems / e2e_steps, "steps": e2e_steps,
               "api": "El.GemmHost -> ElGemmDistHost_d: pinned host [MC,MR] local matrices in, C back in host memory on "
                      "return; C streamed through HBM in column bands, copies overlapped with the SUMMA updates"}
        if serial:
            e2e["unpipelined"] = serial
        # the streamed result is checked too: the host C after T calls against the device product of the same inputs
        try:
            A = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 1)
            B = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 2)
            Ch = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n)
            L.ElDistMatrixLocalFromHost_d(Ch._h, C.c_void_p(pinned[2].data_ptr()), max(lh, 1))
            e2e["parity_NN"] = gemm_parity(El.NORMAL, El.NORMAL, A, B, Ch, 1 + e2e_steps)
            del A, B, Ch
        except Exception as ex:
            e2e["parity_error"] = repr(ex)[:200]
        del pinned
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu:
        cpu, _ = cpu_reference_gemm(n, nb, 1, 1, min(args.cpu_k, n))

    def _guard(fn, name):
        try:
            return fn()
        except Exception as ex:
            import traceback
            traceback.print_exc(file=sys.stderr)
            try:
                L.elb200_sgemm_set_mode(0)
            except Exception:
                pass
            torch.cuda.empty_cache()
            return {"workload": name, "error": repr(ex)[:300]}

    # ---- ZHPDSolve (BASELINE.json configs[3]) ----
    def run_hpd():
        hpd = None
        if not args.no_hpdsolve:
            hn, hr = args.hpd_n, args.hpd_rhs
            El.SetBlocksize(nb)
            Az = El.DistMatrix(np.complex128, El.MC, El.MR, grid, hn, hn).HashFill(1, 11, float(hn))
            Bz = El.DistMatrix(np.complex128, El.MC, El.MR, grid, hn, hr)
            times = []
            for it in range(2):
                Bz.HashFill(0, 13)
                t = timed(lambda: El.HPDSolve(El.LOWER, El.NORMAL, Az, Bz), 1)
                if it > 0:
                    times.append(t)
            hms = sum(times) / len(times)
            hflops = 4.0 * (hn ** 3 / 3.0 + 2.0 * hn * hn * hr)
            hpd = {"workload": f"El::HPDSolve LOWER Complex<double> n={hn} rhs={hr} nb={nb}", "ms": hms,
                   "value": hflops / (hms * 1e-3) / 1e9, "unit": "GFLOP/s (real flops: 4 (n^3/3 + 2 n^2 rhs))",
                   "frac_of_dmma_peak": hflops / (hms * 1e-3) / 1e12 / (dmma_peak_tf * N)}
            Rz = El.DistMatrix(np.complex128, El.MC, El.MR, grid, hn, hr).HashFill(0, 13)
            El.Gemm(El.NORMAL, El.NORMAL, -1.0, Az, Bz, 1.0, Rz)
            hpd["residual"] = El.FrobeniusNorm(Rz) / (hn * 2.0 ** -52 * El.FrobeniusNorm(Az) * El.FrobeniusNorm(Bz))
            hpd["residual_def"] = "||A X - B||_F / (n eps ||A||_F ||X||_F)"
            del Az, Bz, Rz
            torch.cuda.empty_cache()
        return hpd

    # ---- SGEMM SUMMA_Dot (BASELINE.json configs[4]): exact FFMA vs 3xTF32 on tcgen05 ----
    def run_sg():
        sg = None
        if not args.no_sgemm:
            sm_, sk = args.sgemm_mn, args.sgemm_k
            El.SetBlocksize(nb)
            Af = El.DistMatrix(np.float32, El.MC, El.MR, grid, sm_, sk).HashFill(0, 21)
            Bf = El.DistMatrix(np.float32, El.MC, El.MR, grid, sk, sm_).HashFill(0, 22)
            Cf = El.DistMatrix(np.float32, El.MC, El.MR, grid, sm_, sm_)
            # FP64 product of a 64 x 64 corner (host, from gathered views) for the error of each mode
            bs = min(64, sm_)
            vA, vB = El.DistMatrix(np.float32, El.MC, El.MR, grid), El.DistMatrix(np.float32, El.MC, El.MR, grid)
            a_blk = vA.View(Af, 0, 0, bs, sk).ToGlobal().astype(np.float64)
            b_blk = vB.View(Bf, 0, 0, sk, bs).ToGlobal().astype(np.float64)
            ref_blk = a_blk @ b_blk
            den = sk * 2.0 ** -23 * np.linalg.norm(a_blk) * np.linalg.norm(b_blk)
            sflops = 2.0 * sm_ * sm_ * sk
            sg = {"workload": f"El::Gemm NN float m=n={sm_} k={sk} GEMM_SUMMA_DOT (auto-selected, NN.hpp:305)",
                  "unit": "GFLOP/s", "error_def": "||C - C_fp64||_F / (k eps32 ||A||_F ||B||_F) on the leading 64 x 64 block"}
            try:
                for mode, name in ((1, "3xtf32_tcgen05"), (0, "exact_ffma")):
                    L.elb200_sgemm_set_mode(mode)
                    fn = lambda: El.Gemm(El.NORMAL, El.NORMAL, 1.0, Af, Bf, 0.0, Cf)
                    timed(fn, 1)
                    sms = timed(fn, 1 if mode == 0 else 2) / (1 if mode == 0 else 2)
                    vC = El.DistMatrix(np.float32, El.MC, El.MR, grid)
                    got = vC.View(Cf, 0, 0, bs, bs).ToGlobal().astype(np.float64)
                    sg[name] = {"ms": sms, "value": sflops / (sms * 1e-3) / 1e9, "error": float(np.linalg.norm(got - ref_blk) / den),
                                "kernel": int(L.elb200_sgemm_last_kernel())}
            finally:
                L.elb200_sgemm_set_mode(0)
            del Af, Bf, Cf, vA, vB, vC
            torch.cuda.empty_cache()
        return sg

    # ---- configs[1] also names the NT and TN orientations: one timed call each (extras run last, so that
    #      nothing they do can cost the lines above) ----
    def run_orient():
        out = {}
        A = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 1)
        B = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 2)
        Cm = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, n).HashFill(0, 3)
        El.SetBlocksize(nb)
        for name, oa, ob in (("NT", El.NORMAL, El.TRANSPOSE), ("TN", El.TRANSPOSE, El.NORMAL)):
            fn = lambda: El.Gemm(oa, ob, 1.0, A, B, 1.0, Cm, El.GEMM_SUMMA_C)
            Cm.HashFill(0, 3)
            timed(fn, 1)
            oms = timed(fn, 1)
            out[name] = {"ms": oms, "value": flops_step / (oms * 1e-3) / 1e9, "unit": "GFLOP/s"}
            try:
                out[name]["parity"] = gemm_parity(oa, ob, A, B, Cm, 2)
            except Exception as ex:
                out[name]["parity_error"] = repr(ex)[:200]
        del A, B, Cm
        torch.cuda.empty_cache()
        return out

    # ---- the other paths of section 8 that had no performance line: stationary-A / -B SUMMA on a shape that selects
    #      them by the reference's rule (Gemm/NN.hpp:304-313), and the UPPER factorisation ----
    def run_more():
        out = {}
        El.SetBlocksize(nb)
        kk = max(nb, n // 8)
        for name, (mm, nn_) in (("summa_a", (n, kk)), ("summa_b", (kk, n))):
            A = El.DistMatrix(np.float64, El.MC, El.MR, grid, mm, n).HashFill(0, 1)
            B = El.DistMatrix(np.float64, El.MC, El.MR, grid, n, nn_).HashFill(0, 2)
            Cx = El.DistMatrix(np.float64, El.MC, El.MR, grid, mm, nn_).HashFill(0, 3)
            alg = El.GEMM_SUMMA_A if name == "summa_a" else El.GEMM_SUMMA_B
            fn = lambda: El.Gemm(El.NORMAL, El.NORMAL, 1.0, A, B, 1.0, Cx, alg)
            timed(fn, 1)
            t = timed(fn, 1)
            out[name] = {"workload": f"El::Gemm NN double m={mm} n={nn_} k={n} {name.upper()} nb={nb}", "ms": t,
                         "value": 2.0 * mm * nn_ * n / (t * 1e-3) / 1e9, "unit": "GFLOP/s"}
            del A, B, Cx
            torch.cuda.empty_cache()
        if not args.no_potrf:
            pn, pnb = args.potrf_n, args.potrf_nb
            El.SetBlocksize(pnb)
            H = El.DistMatrix(np.float64, El.MC, El.MR, grid, pn, pn)
            ts = []
            for it in range(2):
                H.HashFill(1, 5, float(pn))
                t = timed(lambda: El.Cholesky(El.UPPER, H), 1)
                if it > 0:
                    ts.append(t)
            out["dpotrf_upper"] = {"workload": f"El::Cholesky UPPER double HPD n={pn} nb={pnb}", "ms": ts[0],
                                   "value": (pn ** 3 / 3.0) / (ts[0] * 1e-3) / 1e9, "unit": "GFLOP/s"}
            del H
            torch.cuda.empty_cache()
            El.SetBlocksize(nb)
        # LU with partial pivoting with the reference driver's solve check
        # (tests/lapack_like/LU.cpp: ||A X - B|| after lu::SolveAfter); flops 2 n^3 / 3
        def run_lu():
            ln = n
            A = El.DistMatrix(np.float64, El.MC, El.MR, grid, ln, ln)
            F = El.DistMatrix(np.float64, El.MC, El.MR, grid, ln, ln)
            ts = []
            P = None
            for it in range(2):
                A.HashFill(0, 77)
                El.Copy(A, F)
                P = El.DistPermutation(grid)
                t = timed(lambda: El.LU(F, P), 1)
                if it > 0:
                    ts.append(t)
            B = El.DistMatrix(np.float64, El.MC, El.MR, grid, ln, 16).HashFill(0, 78)
            X = El.DistMatrix(np.float64, El.MC, El.MR, grid, ln, 16)
            El.Copy(B, X)
            El.LUSolveAfter(El.NORMAL, F, X, P)
            El.Gemm(El.NORMAL, El.NORMAL, -1.0, A, X, 1.0, B)
            res = El.FrobeniusNorm(B) / (ln * np.finfo(np.float64).eps * El.FrobeniusNorm(A) * El.FrobeniusNorm(X))
            return {"workload": f"El::LU (partial pivoting) double n={ln} nb={nb}", "ms": ts[0],
                    "value": (2.0 * ln ** 3 / 3.0) / (ts[0] * 1e-3) / 1e9, "unit": "GFLOP/s", "solve_residual": res,
                    "solve_residual_def": "||A X - B||_F / (n eps ||A||_F ||X||_F), X from lu::SolveAfter, 16 rhs"}
        try:
            out["lu_partial"] = run_lu()
        except Exception as ex:   # an extra: it must not cost the lines above
            out["lu_partial"] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()
        return out

    orient = _guard(run_orient, "dgemm_orientations") if not args.no_orient else None
    more = _guard(run_more, "more") if not args.no_orient else None
    hpd = _guard(run_hpd, "zhpdsolve") if not args.no_hpdsolve else None
    sg = _guard(run_sg, "sgemm_dot") if not args.no_sgemm else None

    if rank == 0:
        lh_, lw_ = (n + r - 1) // r, (n + c - 1) // c
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"El::Gemm NN double m=n=k={n} GEMM_SUMMA_C nb={nb} on DistMatrix<double,MC,MR>",
                       "grid": f"{r}x{c}", "alpha": 1.0, "beta": 1.0,
                       "inputs": "counter-hash uniform[-1,1) on global indices, resident in HBM",
                       "l2": "operands (8.6 GB each at n=32768) far exceed the 126 MB L2; no flush needed"},
            "frac_of_dmma_peak": value / 1e3 / (dmma_peak_tf * N),
            "roofline": {"bound": "tensor", "achieved": kernel_tf, "peak": dmma_peak_tf, "unit": "TFLOP/s",
                         "frac": kernel_tf / dmma_peak_tf if dmma_peak_tf else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": 8.0 * (lh_ * nb + nb * lw_ + 2.0 * lh_ * lw_),
                         "kernel": "elb200::gemm_f64_ws_kernel (persistent, TMA-fed DMMA.8x8x4, tile-info ring, L2 red epilogue)",
                         "launches": int(kcount.value), "kernel_ms_per_step": kms.value / args.steps,
                         "kernel_share_of_step": kms.value / ms if ms else None,
                         "flops_per_launch": kflops.value / max(kcount.value, 1),
                         "launch_shape": f"m={lh_} n={lw_} k={nb} (rank-nb update of the local C)",
                         "peak_source": "FP64 DMMA ceiling measured in this run by elb200_dmma_peak "
                                        "(register-resident mma.sync.m8n8k4.f64 loop on all SMs); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "cublas_dgemm_8192_tflops": cublas_tf, "nominal_fp64_tensor_tflops": 40.0},
            "gpu_launches": launches,
            "parity": parity,
            "redist": stats,
            "clocks": clocks,
        }
        if orient:
            line["dgemm_orientations"] = orient
        if more:
            line["other_paths"] = more
        if potrf:
            line["dpotrf"] = potrf
        if hpd:
            line["zhpdsolve"] = hpd
        if sg:
            line["sgemm_dot"] = sg
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
