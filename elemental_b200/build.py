"""In-tree build of libelb200.so (nvcc, sm_100a only).

`python -m elemental_b200.build` or `build()` from `__graft_entry__`.  Objects
land in elemental_b200/_build/, the library in elemental_b200/libelb200.so (both
git-ignored; the .so travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libelb200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nccl_dirs():
    import importlib.util

    spec = importlib.util.find_spec("nvidia")
    inc = lib = None
    if spec and spec.submodule_search_locations:
        for base in spec.submodule_search_locations:
            cand = Path(base) / "nccl"
            if (cand / "include" / "nccl.h").exists():
                inc, lib = cand / "include", cand / "lib"
                break
    if inc is None and Path("/usr/include/nccl.h").exists():
        inc, lib = Path("/usr/include"), Path("/usr/lib/x86_64-linux-gnu")
    if inc is None:
        raise RuntimeError("nccl.h not found")
    return inc, lib


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cpp")))


def _newest_header_mtime():
    hs = list(CSRC.rglob("*.hpp")) + list(CSRC.rglob("*.cuh")) + list((ROOT / "include").rglob("*.h")) \
        + list((ROOT / "include").rglob("*.hpp"))
    return max((h.stat().st_mtime for h in hs), default=0.0)


def _compile(src: Path, obj: Path, nccl_inc: Path, verbose: bool):
    cmd = [_nvcc(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
           "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(nccl_inc),
           "-x", "cu", "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, (r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> Path:
    OBJ.mkdir(exist_ok=True)
    nccl_inc, nccl_lib = _nccl_dirs()
    hdr_m = _newest_header_mtime()
    todo, objs = [], []
    for s in sources():
        o = OBJ / (s.relative_to(CSRC).as_posix().replace("/", "__") + ".o")
        objs.append(o)
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_m):
            todo.append((s, o))
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
            futs = [ex.submit(_compile, s, o, nccl_inc, verbose) for s, o in todo]
            failed = False
            for f in cf.as_completed(futs):
                src, rc, out = f.result()
                if out.strip() and (rc != 0 or verbose):
                    print(f"--- {src.name} ---\n{out}", file=sys.stderr)
                failed |= rc != 0
            if failed:
                raise RuntimeError("nvcc failed")
    if todo or not LIB.exists():
        cmd = [_nvcc(), *ARCH, "-shared", "-o", str(LIB), *map(str, objs),
               "-L", str(nccl_lib), "-l:libnccl.so.2",
               "-Xlinker", f"-rpath={nccl_lib}", "-Xlinker", "--no-undefined", "-lcuda" if False else "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
