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
    hdrs = [os.path.join(ROOT, "pypose_b200", "csrc", h) for h in ("lie_math.cuh", "lie_ops.cuh", "lm_math.cuh", "imu_cov_math.cuh")]
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


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def poseinv_trial(P, X, scale, dmin, dmax, kind=0, delta=1.0):
    P, X = np.ascontiguousarray(P), np.ascontiguousarray(X)
    Pt, sums = np.empty_like(P), np.zeros(4)
    lib().hostmath_poseinv_trial(int(P.dtype == np.float64), _vp(P), _vp(X), _vp(Pt), _vp(sums), ctypes.c_double(scale),
                                 ctypes.c_double(dmin), ctypes.c_double(dmax), int(kind), ctypes.c_double(delta),
                                 ctypes.c_longlong(P.shape[0]))
    return Pt, sums


def pgo_linearize(nodes, Z, ei, ej, kind=0, delta=1.0):
    nodes, Z = np.ascontiguousarray(nodes), np.ascontiguousarray(Z)
    ei, ej = np.ascontiguousarray(ei, dtype=np.int32), np.ascontiguousarray(ej, dtype=np.int32)
    E = Z.shape[0]
    M, u, loss = np.empty((E, 21), nodes.dtype), np.empty((E, 6), nodes.dtype), np.zeros(1)
    lib().hostmath_pgo_linearize(int(nodes.dtype == np.float64), _vp(nodes), _vp(Z), _vp(ei), _vp(ej), _vp(M), _vp(u),
                                 _vp(loss), int(kind), ctypes.c_double(delta), ctypes.c_longlong(E))
    return M, u, loss


def reproj_rows(poses, pts, pix, cidx):
    poses, pts, pix = (np.ascontiguousarray(a) for a in (poses, pts, pix))
    cidx = np.ascontiguousarray(cidx, dtype=np.int32)
    m = pts.shape[0]
    r, J = np.empty((m, 2), poses.dtype), np.empty((m, 2, 6), poses.dtype)
    lib().hostmath_reproj_rows(int(poses.dtype == np.float64), _vp(poses), _vp(pts), _vp(pix), _vp(cidx), _vp(r), _vp(J),
                               ctypes.c_longlong(m))
    return r, J


def spd_inverse(A_packed, k):
    A = np.ascontiguousarray(A_packed)
    out = np.empty_like(A)
    lib().hostmath_spd_inverse(int(A.dtype == np.float64), int(k), _vp(A), _vp(out), ctypes.c_longlong(A.shape[0]))
    return out


def pgo_linearize_w(nodes, Z, ei, ej, W):
    nodes, Z, W = (np.ascontiguousarray(x) for x in (nodes, Z, W.reshape(-1, 36)))
    ei, ej = np.ascontiguousarray(ei, dtype=np.int32), np.ascontiguousarray(ej, dtype=np.int32)
    E = Z.shape[0]
    M, M0 = np.empty((E, 21), nodes.dtype), np.empty((E, 21), nodes.dtype)
    u, u0 = np.empty((E, 6), nodes.dtype), np.empty((E, 6), nodes.dtype)
    lib().hostmath_pgo_linearize_w(int(nodes.dtype == np.float64), _vp(nodes), _vp(Z), _vp(ei), _vp(ej), _vp(W),
                                   ctypes.c_longlong(36 if W.shape[0] == E and E > 1 else 0), _vp(M), _vp(u), _vp(M0), _vp(u0),
                                   ctypes.c_longlong(E))
    return M, u, M0, u0


def ba_rows(poses, points, cidx, pidx):
    poses, points = np.ascontiguousarray(poses), np.ascontiguousarray(points)
    cidx, pidx = np.ascontiguousarray(cidx, dtype=np.int32), np.ascontiguousarray(pidx, dtype=np.int32)
    m = cidx.shape[0]
    Jc, Jp = np.empty((m, 12), poses.dtype), np.empty((m, 6), poses.dtype)
    lib().hostmath_ba_rows(int(poses.dtype == np.float64), _vp(poses), _vp(points), _vp(cidx), _vp(pidx), _vp(Jc), _vp(Jp),
                           ctypes.c_longlong(m))
    return Jc, Jp


def imu_cov(Rk, Rij, a, dt, gcov, acov, init_cov, chunk):
    """One trajectory: Rk, Rij (F,4), a (F,3), dt (F,1), gcov / acov (1|F,3), init_cov (9,9) -> (9,9)."""
    Rk, Rij, a, dt, gcov, acov, init_cov = (np.ascontiguousarray(x) for x in (Rk, Rij, a, dt, gcov, acov, init_cov))
    F = dt.shape[0]
    cov = np.empty((9, 9), Rk.dtype)
    lib().hostmath_imu_cov(int(Rk.dtype == np.float64), _vp(Rk), _vp(Rij), _vp(a), _vp(dt), _vp(gcov), _vp(acov),
                           ctypes.c_longlong(3 if gcov.shape[0] == F and F > 1 else 0), _vp(init_cov), _vp(cov),
                           ctypes.c_longlong(F), ctypes.c_longlong(chunk))
    return cov


def lm_decide(ctl, cur, trial, predicted, failed):
    """csrc/lm_math.cuh lm_decide on the host: ctl = 14 doubles (optim/_lmstep.py layout) -> 16-double state."""
    c = np.ascontiguousarray(ctl, dtype=np.float64)
    st = np.zeros(16)
    lib().hostmath_lm_decide(_vp(c), ctypes.c_double(cur), ctypes.c_double(trial), ctypes.c_double(predicted),
                             ctypes.c_double(failed), _vp(st))
    return st


def reproj2_rows(nodes, pts, pix, ia, ib, intr):
    nodes, pts, pix = (np.ascontiguousarray(a) for a in (nodes, pts, pix))
    ia, ib = np.ascontiguousarray(ia, dtype=np.int32), np.ascontiguousarray(ib, dtype=np.int32)
    k = np.ascontiguousarray(intr, dtype=np.float64)
    m = pts.shape[0]
    r, J = np.empty((m, 2), nodes.dtype), np.empty((m, 2, 6), nodes.dtype)
    lib().hostmath_reproj2_rows(int(nodes.dtype == np.float64), _vp(nodes), _vp(pts), _vp(pix), _vp(ia), _vp(ib), _vp(k), _vp(r),
                                _vp(J), ctypes.c_longlong(m))
    return r, J
