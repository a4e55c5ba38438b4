"""Per-warp phase clocks of the TMA DGEMM kernels (diagnostic builds, elb200_dgemm_set_debug_flags(1024)):
where the warps of one CTA spend their SM clocks -- producer work, waiting for a full stage, fragment loads + DMMAs,
epilogue -- for the rank-nb update shape and a long-k shape; cfg 3 = third-generation kernel (gemm_f64_ws.cu),
cfg 4 = second generation (gemm_f64_tma.cu).  usage: python scripts/gpu_dgemm_phase_clocks.py [cfgs...]"""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
nsm = torch.cuda.get_device_properties(0).multi_processor_count
cfgs = [int(x) for x in sys.argv[1:]] or [4, 3]
def run(cfg, m, n, k, flags, ta="N", tb="N", alpha=1.0, beta=1.0, quiet=False):
    L.elb200_dgemm_set_config(cfg)
    ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
    B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    prof = torch.zeros(nsm * 16 * 8, dtype=torch.int64, device=dev)
    L.elb200_dgemm_set_profile_buffer(C.c_void_p(prof.data_ptr()))
    L.elb200_dgemm_set_debug_flags(flags)
    fn = lambda: check(L.elb200_dgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(alpha), C.c_void_p(A.data_ptr()), G.i64(ar),
                                      C.c_void_p(B.data_ptr()), G.i64(br), C.c_double(beta), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    L.elb200_dgemm_set_debug_flags(0)
    L.elb200_dgemm_set_profile_buffer(None)
    L.elb200_dgemm_set_config(0)
    p = prof.view(nsm, 16, 8).double().cpu()
    print(f"cfg {cfg} {ta}{tb} {m}x{n}x{k} alpha {alpha} beta {beta} flags {flags}: {2*m*n*k/best/1e9:.2f} TF/s  ({best:.3f} ms)", flush=True)
    if flags & 1024:
        tot = p[:, :, 4].mean().item()
        tiles = p[:, :, 5].mean().item()
        print(f"  kernel clocks per warp (mean) {tot:.0f}; tiles per warp {tiles:.1f}; DMMA issue clocks of a sub-partition (4 warps x tiles x k/4 x 16 DMMA x 16 clk) = {4*tiles*k/4*16*16:.0f} = {100*4*tiles*k/4*16*16/tot:.1f}%")
        names = ["producer", "wait_full", "lds+dmma", "epilogue"]
        for w in range(16):
            row = "  ".join(f"{names[i]} {100*p[:, w, i].mean().item()/tot:5.1f}%" for i in range(4))
            rest = 100 * (1 - sum(p[:, w, i].mean().item() for i in range(4)) / tot)
            print(f"  warp {w:2d}: {row}  other {rest:5.1f}%   epilogue clocks/tile {p[:, w, 3].mean().item()/max(tiles,1):.0f}  lds+dmma clocks/stage {p[:, w, 2].mean().item()/max(tiles*((k+15)//16),1):.0f}")
    del A, B, Cm
for cfg in [3]:
    for (m, n, k) in [(32768, 32768, 128), (16384, 16384, 256), (16384, 8192, 128), (8192, 8192, 8192)]:
        run(cfg, m, n, k, 0)
    run(cfg, 32768, 32768, 128, 0, alpha=-1.0)
    run(cfg, 32768, 32768, 128, 0, alpha=3.0)
    run(cfg, 32768, 32768, 128, 0, beta=0.0)
    for ta, tb in (("N", "T"), ("T", "N"), ("T", "T")):
        run(cfg, 32768, 32768, 128, 0, ta, tb)
print("3-D tensor maps used by the last third-generation launch (bit 0 A, bit 1 B):", L.elb200_dgemm_ws_last_maps())
run(3, 32768, 32768, 128, 0); print("maps NN:", L.elb200_dgemm_ws_last_maps())
run(3, 32768, 32768, 128, 1024)
run(3, 8192, 8192, 8192, 1024)
# bisect (diagnostic build only; results are wrong on purpose): flags bits 16.. = 1 no epilogue, 2 no TMA, 4 no fragment loads
for dbg in ():
    run(3, 32768, 32768, 128, 1024 | (dbg << 16))
for dbg in ():
    run(3, 8192, 8192, 8192, 1024 | (dbg << 16))
