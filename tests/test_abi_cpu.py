"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol that include/*.h declares; compute calls fail loudly (no CPU fallback); host-side index
math and the SUMMA selector match the reference's definitions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from elemental_b200._lib import LIB_PATH, lib
from oracle import elemental_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("elb200_blas.h", "elb200_level1.h", "elb200_plan.h", "elb200_El.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        if h == "elb200_El.h":
            # expand the per-type macro block
            m = re.search(r"#define ELB200_DECLARE_TYPE\(SUF, SCALAR, REAL\)(.*?)\n\nELB200_DECLARE_TYPE", text, flags=re.S)
            body = m.group(1).replace("\\\n", "\n")
            for suf in "sdcz":
                for fn in re.findall(r"\b(El\w+)_##SUF\s*\(", body):
                    names.add(f"{fn}_{suf}")
            text = text.replace(m.group(0), "")
        for fn in re.findall(r"\b((?:elb200_|El)[A-Za-z0-9_]+|[sdcz](?:gemm|trsm|syrk|herk)_)\s*\(", text):
            if not fn.startswith("ELB200"):
                names.add(fn)
    return sorted(names)


def test_library_loads_and_exports_every_declared_symbol():
    assert LIB_PATH.exists(), "build the library first: python -m elemental_b200.build"
    L = lib()
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert len(_declared_symbols()) > 250


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib()
    L.elb200_last_error.restype = C.c_char_p
    assert L.elb200_device_check() != 0
    assert b"no CPU fallback" in L.elb200_last_error() or b"CUDA" in L.elb200_last_error()
    from elemental_b200 import api
    with pytest.raises(Exception):
        api.Initialize()


def test_product_path_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "elemental_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/generator.py", "").replace("oracle/elemental_oracle.py", ""), f


def test_index_math_matches_oracle():
    L = lib()
    L.elb200_shift.restype = C.c_int64
    L.elb200_length.restype = C.c_int64
    for stride in (1, 2, 3, 4, 8):
        for align in range(stride):
            for rank in range(stride):
                sh = L.elb200_shift(C.c_int64(rank), C.c_int64(align), C.c_int64(stride))
                assert sh == O.shift(rank, align, stride)
                for n in (0, 1, 5, 17, 64):
                    assert L.elb200_length(C.c_int64(n), C.c_int64(sh), C.c_int64(stride)) == O.length(n, sh, stride)
    names = {0: "MC", 2: "MR", 3: "VC", 4: "VR", 5: "STAR"}
    for r, c in [(1, 1), (2, 4), (3, 2)]:
        for d, nm in names.items():
            assert L.elb200_dist_stride(d, r, c) == O.dist_stride(nm, r, c)
            for i in range(r):
                for j in range(c):
                    assert L.elb200_dist_rank(d, r, c, i, j) == O.dist_rank(nm, i, j, r, c)


def test_gemm_selector_matches_reference_rule():
    """Gemm/NN.hpp:304-313: Dot if 10m<=k and 10n<=k; B if m<=n and 2m<=k; A if n<=m and 2n<=k; else C."""
    L = lib()
    rng = np.random.default_rng(0)
    for _ in range(500):
        m, n, k = (int(x) for x in rng.integers(1, 5000, 3))
        assert L.elb200_gemm_default_algorithm(C.c_int64(m), C.c_int64(n), C.c_int64(k)) == O.gemm_select(m, n, k)
    assert L.elb200_gemm_default_algorithm(C.c_int64(8192), C.c_int64(8192), C.c_int64(262144)) == O.GEMM_SUMMA_DOT
    assert L.elb200_gemm_default_algorithm(C.c_int64(32768), C.c_int64(32768), C.c_int64(32768)) == O.GEMM_SUMMA_C


def test_tf32_tile_raster_is_a_bijection():
    """The 3xTF32 kernel's producer and epilogue warps map a linear tile index to (tile row, tile column) with
    elb200_tf32_tile_coords (host copy of the device function): every tile exactly once, ragged last band included."""
    import ctypes as C
    from elemental_b200._lib import lib
    L = lib()
    L.elb200_tf32_tile_coords.restype = None
    tm, tn = C.c_int64(), C.c_int64()
    for tilesM, tilesN in [(1, 1), (1, 40), (40, 1), (7, 16), (7, 17), (64, 64), (16, 31), (3, 33), (5, 48)]:
        seen = set()
        for t in range(tilesM * tilesN):
            L.elb200_tf32_tile_coords(C.c_int64(t), C.c_int64(tilesM), C.c_int64(tilesN), C.byref(tm), C.byref(tn))
            assert 0 <= tm.value < tilesM and 0 <= tn.value < tilesN, (tilesM, tilesN, t, tm.value, tn.value)
            seen.add((tm.value, tn.value))
        assert len(seen) == tilesM * tilesN, (tilesM, tilesN)
    # consecutive tiles stay within a band of 16 tile columns (the L2 reuse the raster is for)
    cols = set()
    for t in range(148):
        L.elb200_tf32_tile_coords(C.c_int64(t), C.c_int64(64), C.c_int64(64), C.byref(tm), C.byref(tn))
        cols.add(tn.value)
    assert len(cols) <= 16
