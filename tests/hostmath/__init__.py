"""TEST-ONLY host build of the device math headers (see hostmath.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libhostmath.so")
_lib = None


def build():
    src = os.path.join(HERE, "hostmath.cpp")
    hdrs = [os.path.join(ROOT, "pypose_b200", "csrc", h) for h in ("lie_math.cuh", "lie_ops.cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-x", "c++", "-shared", "-fPIC",
                               "-I", os.path.join(ROOT, "pypose_b200", "csrc"), src, "-o", SO])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostmath_run.restype = ctypes.c_int
    return _lib


def run(group, op, ins, out_widths):
    """ins: list of (N, d) float32/float64 arrays -> list of (N, w) outputs."""
    ins = [np.ascontiguousarray(a) for a in ins]
    dt = ins[0].dtype
    n = ins[0].shape[0]
    outs = [np.empty((n, w), dtype=dt) for w in out_widths]
    PI = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    PO = (ctypes.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
    rc = lib().hostmath_run(group.encode(), op.encode(), int(dt == np.float64), PI, PO, ctypes.c_longlong(n))
    if rc != 0:
        raise RuntimeError(f"hostmath_run({group},{op}) -> {rc}")
    return outs
