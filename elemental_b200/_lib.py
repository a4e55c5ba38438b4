"""ctypes loader for libelb200.so.

The product path has no CPU fallback: if the library is missing or no sm_100
device is visible, the calls raise.  The checker package (the CPU oracle) is never imported here.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libelb200.so"

_lib = None


class Elb200Error(RuntimeError):
    pass


class c32(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class c64(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library; raise loudly when absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise Elb200Error(
                f"{LIB_PATH} is missing: build it with `python -m elemental_b200.build` "
                "(there is no CPU fallback for this path)")
        # torch first, so that libnccl.so.2 / CUDA libs resolve to the copies it ships
        import torch  # noqa: F401

        _lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
        _lib.elb200_last_error.restype = C.c_char_p
        _lib.elb200_get_stream.restype = C.c_void_p
    return _lib


def check(rc: int, what: str = "elb200 call") -> None:
    if rc != 0:
        msg = lib().elb200_last_error()
        raise Elb200Error(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def require_device() -> None:
    check(lib().elb200_device_check(), "elb200_device_check")
