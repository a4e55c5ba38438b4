"""Dry run of bench.py's own arm on the CPU: the device API (elemental_b200.api, the ctypes library, torch.cuda)
is replaced by recording fakes, so that the control flow, the order of the sections and the assembly of the JSON line
are executed for real.  This is a test of bench.py's HOST logic only (a NameError or a dropped key here would cost the
round's measurement); it makes no parity or performance claim and nothing in it touches the product path."""
import ctypes as C
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeDM:
    def __init__(self, dtype=np.float64, colDist=0, rowDist=2, grid=None, height=0, width=0):
        self.dtype, self.h, self.w, self._h = np.dtype(dtype), height, width, C.c_void_p(1)

    def HashFill(self, *a):
        return self

    def LocalHeight(self): return self.h
    def LocalWidth(self): return self.w

    def View(self, parent, i, j, h, w):
        self.h, self.w = h, w
        return self

    def ToGlobal(self):
        return np.ones((self.h, self.w), dtype=self.dtype)


class _FakeGrid:
    def __init__(self, height=0): pass
    def Height(self): return 1
    def Width(self): return 1


def _fake_api(calls):
    m = types.ModuleType("elemental_b200.api")
    m.DistMatrix, m.Grid = _FakeDM, _FakeGrid
    for i, name in enumerate(("MC", "MD", "MR", "VC", "VR", "STAR")):
        setattr(m, name, i)
    m.NORMAL, m.TRANSPOSE, m.ADJOINT, m.LOWER, m.UPPER = 0, 1, 2, 0, 1
    m.GEMM_DEFAULT, m.GEMM_SUMMA_A, m.GEMM_SUMMA_B, m.GEMM_SUMMA_C, m.GEMM_SUMMA_DOT = range(5)

    def rec(name, ret=None):
        def f(*a, **k):
            calls.append(name)
            return ret
        return f
    for name in ("SetBlocksize", "Gemm", "GemmHost", "Cholesky", "CholeskySolveAfter", "HPDSolve"):
        setattr(m, name, rec(name))
    m.FrobeniusNorm = rec("FrobeniusNorm", 1.0)
    m.RedistStats = rec("RedistStats", {"copies": 0})
    return m


class _FakeLib:
    def __getattr__(self, name):
        def f(*a):
            if name == "elb200_dmma_peak":
                a[1]._obj.value = 37.0e12
            if name == "elb200_gemm_profile_read":
                a[0]._obj.value, a[1]._obj.value, a[2]._obj.value = 10.0, 4, 4.0e11
            if name == "elb200_sgemm_last_kernel":
                return 2
            return 0
        f.restype = None
        self.__dict__[name] = f
        return f


def test_own_arm_control_flow_and_json_line(monkeypatch):
    import torch
    calls = []
    fake_api = _fake_api(calls)
    fake_lib = types.ModuleType("elemental_b200._lib")
    fake_lib.lib = lambda: _FakeLib()
    pkg = types.ModuleType("elemental_b200")
    pkg.api, pkg._lib = fake_api, fake_lib
    monkeypatch.setitem(sys.modules, "elemental_b200", pkg)
    monkeypatch.setitem(sys.modules, "elemental_b200.api", fake_api)
    monkeypatch.setitem(sys.modules, "elemental_b200._lib", fake_lib)

    class Ev:
        def __init__(self, enable_timing=False): pass
        def record(self): pass
        def elapsed_time(self, other): return 5.0
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    real_tensor, real_empty = torch.tensor, torch.empty
    monkeypatch.setattr(torch, "randn", lambda *s, **k: real_empty(2, 2, dtype=torch.float64))
    monkeypatch.setattr(torch, "matmul", lambda a, b: a)
    monkeypatch.setattr(torch, "tensor", lambda data, device=None, dtype=None: real_tensor(data, dtype=dtype))
    monkeypatch.setattr(torch, "empty", lambda shape, dtype=None, pin_memory=False: real_empty(shape, dtype=dtype))
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1", "--n", "256", "--potrf-n", "256",
                                      "--hpd-n", "128", "--hpd-rhs", "32", "--sgemm-mn", "128", "--sgemm-k", "512",
                                      "--no-cpu"])
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965, "sm_max_mhz": 1965, "reasons": [], "samples": 1})
    out = io.StringIO()
    with redirect_stdout(out):
        bench.main()
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "gpu_launches", "clocks", "e2e", "dpotrf", "parity",
                "zhpdsolve", "sgemm_dot", "dgemm_orientations"):
        assert key in d, key
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in d["roofline"], key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["config"]["workload"]
    assert "error" not in d["zhpdsolve"] and "error" not in d["sgemm_dot"] and "error" not in d["dgemm_orientations"]
    assert set(d["dgemm_orientations"]) == {"NT", "TN"}
    assert {"3xtf32_tcgen05", "exact_ffma"} <= set(d["sgemm_dot"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert "GemmHost" in calls and "NN" in d["parity"] and "parity_NN" in d["e2e"]
    assert all("parity" in v for v in d["dgemm_orientations"].values())
    # the headline sections run before the extras
    first_extra = min(i for i, c in enumerate(calls) if c == "HPDSolve")
    assert "Cholesky" in calls[:first_extra]
