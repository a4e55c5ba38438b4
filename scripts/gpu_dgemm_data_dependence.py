"""Is the FP64 tensor rate data dependent?  The same DGEMM launch (rank-128 update and 8192^3) on operands that are
all ones, small integers, or uniform random; SM clock sampled while it runs."""
import ctypes as C, subprocess, sys, threading, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
def clocks():
    out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    return out
def run(m, n, k, kind, cfg=3, reps=40):
    L.elb200_dgemm_set_config(cfg)
    A = torch.empty(k, m, dtype=torch.float64, device=dev); B = torch.empty(n, k, dtype=torch.float64, device=dev)
    if kind == "ones": A.fill_(1.0); B.fill_(1.0)
    elif kind == "ints": A.copy_(torch.randint(-3, 4, A.shape, device=dev).double()); B.copy_(torch.randint(-3, 4, B.shape, device=dev).double())
    elif kind == "zeros": A.zero_(); B.zero_()
    else: A.uniform_(-1, 1); B.uniform_(-1, 1)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    fn = lambda: check(L.elb200_dgemm(G.ch("N"), G.ch("N"), G.i64(m), G.i64(n), G.i64(k), C.c_double(1.0), C.c_void_p(A.data_ptr()), G.i64(m),
                                      C.c_void_p(B.data_ptr()), G.i64(k), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    samples = []
    stop = [False]
    th = threading.Thread(target=lambda: [samples.append(clocks()) or time.sleep(0.05) for _ in iter(lambda: stop[0], True)])
    th.start()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    ms = e0.elapsed_time(e1) / reps
    print(f"cfg {cfg} NN {m}x{n}x{k} {kind:7s}: {2*m*n*k/ms/1e9:.2f} TF/s  ({ms:.3f} ms)  clocks/power samples: {samples[len(samples)//2:len(samples)//2+3]}", flush=True)
    L.elb200_dgemm_set_config(0)
    del A, B, Cm
for kind in ("random", "ones", "ints", "zeros", "random"):
    run(32768, 32768, 128, kind)
for kind in ("random", "ones", "zeros"):
    run(8192, 8192, 8192, kind, reps=10)
for kind in ("random", "ones"):
    run(32768, 32768, 128, kind, cfg=4)
