"""ctypes binding of the PLAIN layer of include/suitesparse_b200.h (ssb200_*): raw arrays in, raw arrays out.
The library is required — there is no CPU fallback; a missing .so or a missing GPU raises."""
from __future__ import annotations
import ctypes as C
import numpy as np
from .cholmod_host import load_b200

c_long = C.c_int64


class Stats(C.Structure):
    _fields_ = [("nsuper", c_long), ("nlevels", c_long), ("nupdates", c_long),
                ("kernel_launches", c_long), ("kernel_launches_total", c_long),
                ("flops_update", C.c_double), ("flops_potrf", C.c_double), ("flops_trsm", C.c_double),
                ("ms_total", C.c_double), ("ms_assemble", C.c_double), ("ms_update", C.c_double),
                ("ms_factor", C.c_double), ("ms_d2h", C.c_double), ("ms_h2d", C.c_double),
                ("bytes_update_panel", C.c_double), ("bytes_update_scatter", C.c_double),
                ("device_bytes", c_long),
                ("ms_kind", C.c_double * 6), ("flops_kind", C.c_double * 6), ("launches_kind", c_long * 6),
                ("d2h_staged", c_long)]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        for k in ("ms_kind", "flops_kind", "launches_kind"):
            d[k] = list(d[k])
        return d


def _lib():
    L = load_b200()
    if not getattr(L, "_ssb_typed", False):
        L.ssb200_plan_create.restype = C.c_void_p
        L.ssb200_plan_create.argtypes = [c_long, c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ssb200_plan_destroy.argtypes = [C.c_void_p]
        L.ssb200_factorize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_long,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.POINTER(c_long)]
        L.ssb200_upload_A.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_long,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ssb200_factorize_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(c_long)]
        L.ssb200_download_L.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_upload_L.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, c_long, c_long]
        L.ssb200_solve_resident.argtypes = [C.c_void_p, C.c_int, C.c_void_p, c_long, c_long]
        L.ssb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.ssb200_last_error.restype = C.c_char_p
        L.ssb200_version.restype = C.c_char_p
        L.ssb200_device_Lx.restype = C.c_void_p
        L.ssb200_device_Lx.argtypes = [C.c_void_p]
        L.ssb200_xsize.restype = c_long
        L.ssb200_xsize.argtypes = [C.c_void_p]
        L.ssb200_debug_relmap.restype = c_long
        L.ssb200_debug_relmap.argtypes = [C.c_void_p, C.c_void_p, c_long]
        L.ssb200_factor_diag.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_plan_of_factor.restype = C.c_void_p
        L.ssb200_plan_of_factor.argtypes = [C.c_void_p]
        L._ssb_typed = True
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SsbError(RuntimeError):
    pass


class Plan:
    """Device plan + device-resident factor for one symbolic supernodal structure (super/pi/px/s)."""

    def __init__(self, n, super_, pi, px, s, device: int = -1, handle=None):
        self.lib = _lib()
        self.n = int(n)
        self.owned = handle is None
        if handle is None:
            self.super, self.pi, self.px, self.s = _i64(super_), _i64(pi), _i64(px), _i64(s)
            self.nsuper = self.super.size - 1
            handle = self.lib.ssb200_plan_create(self.n, self.nsuper, _ptr(self.super), _ptr(self.pi), _ptr(self.px),
                                                 _ptr(self.s), device)
            if not handle:
                raise SsbError(self.lib.ssb200_last_error().decode())
        self.h = C.c_void_p(handle)
        self.xsize = int(self.lib.ssb200_xsize(self.h))

    def _check(self, rc):
        if rc < 0:
            raise SsbError(f"status {rc}: {self.lib.ssb200_last_error().decode()}")
        return rc

    def close(self):
        if self.h and self.owned:
            self.lib.ssb200_plan_destroy(self.h)
        self.h = None

    def upload_A(self, A_lower, F=None):
        Ap, Ai, Ax = _i64(A_lower.indptr), _i64(A_lower.indices), np.ascontiguousarray(A_lower.data, dtype=np.float64)
        self._keepA = (Ap, Ai, Ax)
        if F is None:
            return self._check(self.lib.ssb200_upload_A(self.h, -1, _ptr(Ap), _ptr(Ai), None, _ptr(Ax), A_lower.shape[1],
                                                        None, None, None, None))
        Fp, Fi, Fx = _i64(F.indptr), _i64(F.indices), np.ascontiguousarray(F.data, dtype=np.float64)
        return self._check(self.lib.ssb200_upload_A(self.h, 0, _ptr(Ap), _ptr(Ai), None, _ptr(Ax), A_lower.shape[1],
                                                    _ptr(Fp), _ptr(Fi), None, _ptr(Fx)))

    def factorize_resident(self, beta: float = 0.0, quick_return: bool = False):
        b = (C.c_double * 2)(beta, 0.0)
        minor = c_long(0)
        st = self._check(self.lib.ssb200_factorize_resident(self.h, b, 1 if quick_return else 0, C.byref(minor)))
        return st, minor.value

    def download_L(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.xsize, dtype=np.float64)
        self._check(self.lib.ssb200_download_L(self.h, _ptr(out)))
        return out

    def factorize(self, A_lower, beta: float = 0.0, quick_return: bool = False, F=None):
        """One call = upload A, factorize on the device, download L.  Returns (status, minor, Lx)."""
        self.upload_A(A_lower, F)
        st, minor = self.factorize_resident(beta, quick_return)
        return st, minor, self.download_L()

    def solve(self, X: np.ndarray, which: int = 2) -> np.ndarray:
        X = np.array(X, dtype=np.float64, order="F", copy=True)
        X2 = X.reshape(X.shape[0], -1, order="F")
        self._check(self.lib.ssb200_solve(self.h, which, _ptr(X2), X2.shape[1], X2.shape[0]))
        return X

    def solve_resident(self, dX_ptr: int, nrhs: int, ldx: int, which: int = 2):
        self._check(self.lib.ssb200_solve_resident(self.h, which, C.c_void_p(dX_ptr), nrhs, ldx))

    def set_lookahead(self, on: bool):
        self.lib.ssb200_set_lookahead.argtypes = [C.c_void_p, C.c_int]
        self._check(self.lib.ssb200_set_lookahead(self.h, 1 if on else 0))

    def factor_diag(self) -> np.ndarray:
        d = np.empty(self.n, dtype=np.float64)
        self._check(self.lib.ssb200_factor_diag(self.h, _ptr(d)))
        return d

    def stats(self) -> dict:
        st = Stats()
        self.lib.ssb200_get_stats(self.h, C.byref(st))
        return st.asdict()

    def relmap(self) -> np.ndarray:
        nrel = self.lib.ssb200_debug_relmap(self.h, None, 0)
        out = np.empty(max(nrel, 1), dtype=np.int32)
        self.lib.ssb200_debug_relmap(self.h, _ptr(out), nrel)
        return out[:nrel]


class MultiGpu:
    """One factorization over several GPUs of this process (ssb200_mg_*): distributed storage, pulls over NVLink."""

    def __init__(self, n, super_, pi, px, s, devices=None, ndev: int | None = None, handle=None):
        self.lib = L = _lib()
        L.ssb200_mg_create.restype = C.c_void_p
        L.ssb200_mg_create.argtypes = [c_long, c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ssb200_mg_destroy.argtypes = [C.c_void_p]
        L.ssb200_mg_pin_host.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_mg_factorize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_long,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(c_long)]
        L.ssb200_mg_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, c_long, c_long]
        L.ssb200_mg_upload_L.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_mg_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.n = int(n)
        self.owned = handle is None
        if handle is None:
            self.super, self.pi, self.px, self.s = _i64(super_), _i64(pi), _i64(px), _i64(s)
            if devices is None:
                devices = list(range(ndev if ndev else L.ssb200_device_count()))
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            handle = L.ssb200_mg_create(self.n, self.super.size - 1, _ptr(self.super), _ptr(self.pi), _ptr(self.px), _ptr(self.s), len(dv), _ptr(dv))
            if not handle:
                raise SsbError(L.ssb200_last_error().decode())
            self.ndev = len(dv); self.xsize = int(self.px[-1])
        self.h = C.c_void_p(handle)

    def _check(self, rc):
        if rc < 0:
            raise SsbError(f"status {rc}: {self.lib.ssb200_last_error().decode()}")
        return rc

    def pin_host(self, Lx: np.ndarray):
        return self.lib.ssb200_mg_pin_host(self.h, _ptr(Lx))

    def factorize(self, A_lower, beta: float = 0.0, Lx_host: np.ndarray | None = None, F=None):
        """Returns (status, minor).  Lx_host (xsize doubles) receives the factor when given."""
        Ap, Ai, Ax = _i64(A_lower.indptr), _i64(A_lower.indices), np.ascontiguousarray(A_lower.data, dtype=np.float64)
        b = (C.c_double * 2)(beta, 0.0)
        minor = c_long(0)
        if F is None:
            st = self.lib.ssb200_mg_factorize(self.h, -1, _ptr(Ap), _ptr(Ai), None, _ptr(Ax), A_lower.shape[1], None, None, None, None,
                                              b, _ptr(Lx_host), C.byref(minor))
        else:
            Fp, Fi, Fx = _i64(F.indptr), _i64(F.indices), np.ascontiguousarray(F.data, dtype=np.float64)
            st = self.lib.ssb200_mg_factorize(self.h, 0, _ptr(Ap), _ptr(Ai), None, _ptr(Ax), A_lower.shape[1], _ptr(Fp), _ptr(Fi), None, _ptr(Fx),
                                              b, _ptr(Lx_host), C.byref(minor))
        return self._check(st), minor.value

    def factorize_resident(self, beta: float = 0.0):
        """The matrix uploaded by the previous factorize() call, no host copy of L."""
        b = (C.c_double * 2)(beta, 0.0)
        minor = c_long(0)
        st = self.lib.ssb200_mg_factorize(self.h, -1, None, None, None, None, 0, None, None, None, None, b, None, C.byref(minor))
        return self._check(st), minor.value

    def launches(self) -> int:
        self.lib.ssb200_mg_launches.restype = c_long; self.lib.ssb200_mg_launches.argtypes = [C.c_void_p]
        return int(self.lib.ssb200_mg_launches(self.h))

    def solve(self, X: np.ndarray, which: int = 2) -> np.ndarray:
        X = np.array(X, dtype=np.float64, order="F", copy=True)
        X2 = X.reshape(X.shape[0], -1, order="F")
        self._check(self.lib.ssb200_mg_solve(self.h, which, _ptr(X2), X2.shape[1], X2.shape[0]))
        return X

    def upload_L(self, Lx: np.ndarray):
        self._check(self.lib.ssb200_mg_upload_L(self.h, _ptr(np.ascontiguousarray(Lx, dtype=np.float64))))

    def info(self) -> dict:
        out = np.zeros(3 + 2 * 16)
        N = self.lib.ssb200_mg_info(self.h, _ptr(out), out.size)
        return dict(ndev=N, ms_factorize=out[0], ms_solve=out[1], nvlink_bytes=out[2], device_bytes=out[3:3 + N].tolist(),
                    rank_flops=out[3 + N:3 + 2 * N].tolist(), ms_device=out[3 + 2 * N])

    def trace(self):
        """(times[ndev, nsteps+1] in ms, steps[nsteps, 4] = src, wait_remote, cnt, next_owner); needs SSB200_MG_TRACE=1."""
        L = self.lib
        L.ssb200_mg_trace.restype = c_long
        L.ssb200_mg_trace.argtypes = [C.c_void_p, C.c_void_p, c_long, C.c_void_p, c_long]
        ns = L.ssb200_mg_trace(self.h, None, 0, None, 0)
        N = self.info()["ndev"]
        t = np.zeros((N, ns + 1), dtype=np.float32); st = np.zeros((ns, 4), dtype=np.int64)
        L.ssb200_mg_trace(self.h, _ptr(t), t.size, _ptr(st), st.size)
        return t, st

    def close(self):
        if self.h and self.owned:
            self.lib.ssb200_mg_destroy(self.h)
        self.h = None


def plan_of_factor(Lp) -> Plan | None:
    """The plan the drop-in layer cached for a cholmod_factor (statistics of the interposed path)."""
    h = _lib().ssb200_plan_of_factor(Lp)
    if not h:
        return None
    return Plan(Lp.contents.n, None, None, None, None, handle=h)
