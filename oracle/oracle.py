"""ctypes wrapper of oracle/ssb_oracle.c — TEST INFRASTRUCTURE (checker / CPU baseline only).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "_ref", "libssb_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libcholmod_ref.so")
HOST_LIB = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "libcholmod.so")   # the application's host library


def build(ref: bool = True):
    """Compile the C restatement, and the reference build when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir(os.environ.get("SSB200_REFERENCE", "/root/reference")):
        if not os.path.exists(REF_LIB):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        if not os.path.exists(HOST_LIB):
            subprocess.check_call(["make", "-s", "-C", HERE, "host"])
        if not os.path.exists(os.path.join(HERE, "_ref", "cholmod_l_demo")):
            subprocess.check_call(["make", "-s", "-C", HERE, "demo"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_LIB):
            build(ref=False)
        _lib = C.CDLL(PORT_LIB)
        _lib.ssbo_enumerate_updates.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def enumerate_updates(n, super_, pi, s):
    super_, pi, s = _i64(super_), _i64(pi), _i64(s)
    nsuper = super_.size - 1
    L = lib()
    mc = C.c_int64(0)
    cnt = L.ssbo_enumerate_updates(C.c_int64(n), C.c_int64(nsuper), _p(super_), _p(pi), _p(s), None, None, None, None, None, C.byref(mc))
    arrs = [np.empty(cnt, dtype=np.int64) for _ in range(5)]
    L.ssbo_enumerate_updates(C.c_int64(n), C.c_int64(nsuper), _p(super_), _p(pi), _p(s), *[_p(a) for a in arrs], C.byref(mc))
    return dict(d=arrs[0], s=arrs[1], p0=arrs[2], ndrow1=arrs[3], ndrow2=arrs[4], maxcsize=mc.value)


def factorize(n, super_, pi, px, s, A_lower, beta=0.0, quick_return=False, F=None):
    """A_lower: scipy CSC holding (at least) the lower triangle of the permuted matrix (stype<0), or the
    unsymmetric A with F=A' (stype==0).  Returns (status, minor, Lx)."""
    super_, pi, px, s = _i64(super_), _i64(pi), _i64(px), _i64(s)
    nsuper = super_.size - 1
    Ap, Ai, Ax = _i64(A_lower.indptr), _i64(A_lower.indices), np.ascontiguousarray(A_lower.data, dtype=np.float64)
    if F is not None:
        Fp, Fi, Fx = _i64(F.indptr), _i64(F.indices), np.ascontiguousarray(F.data, dtype=np.float64)
        stype = 0
    else:
        Fp = Fi = Fx = None
        stype = -1
    Lx = np.zeros(int(px[nsuper]), dtype=np.float64)
    minor = C.c_int64(0)
    b = (C.c_double * 2)(beta, 0.0)
    st = lib().ssbo_factorize(C.c_int64(n), C.c_int64(nsuper), _p(super_), _p(pi), _p(px), _p(s), C.c_int(stype),
                              _p(Ap), _p(Ai), None, _p(Ax), _p(Fp), _p(Fi), None, _p(Fx), b,
                              C.c_int(1 if quick_return else 0), _p(Lx), C.byref(minor))
    return st, minor.value, Lx


def lsolve(super_, pi, px, s, Lx, X, transpose=False):
    super_, pi, px, s = _i64(super_), _i64(pi), _i64(px), _i64(s)
    X = np.array(X, dtype=np.float64, order="F", copy=True)
    X2 = X.reshape(X.shape[0], -1, order="F")
    f = lib().ssbo_ltsolve if transpose else lib().ssbo_lsolve
    f(C.c_int64(super_.size - 1), _p(super_), _p(pi), _p(px), _p(s), _p(np.ascontiguousarray(Lx)), _p(X2),
      C.c_int64(X2.shape[1]), C.c_int64(X2.shape[0]))
    return X
