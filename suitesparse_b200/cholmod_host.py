"""ctypes view of the CHOLMOD C API — the *application side* of the drop-in boundary.

A CHOLMOD user calls cholmod_l_start / analyze / factorize / solve on a host libcholmod
(CHOLMOD/Include/cholmod_cholesky.h:89,113,160,216).  This module is that user, in Python: it loads
  1. libsuitesparse_b200.so with RTLD_GLOBAL (optional: the GPU arm), which defines
     cholmod_l_super_numeric / _lsolve / _ltsolve, and then
  2. the host libcholmod (any build of the reference; on this box baseline/_ref/libcholmod.so),
so the host library's PLT calls at cholmod_factorize.c:265 and cholmod_solve.c:1568-1577 bind to the
B200 implementation — the same interposition LD_PRELOAD gives a C program.
The struct layouts restate include/suitesparse_b200.h (which restates cholmod_core.h).
"""
from __future__ import annotations
import ctypes as C
import os
import numpy as np

c_long = C.c_int64
MAXMETHODS = 9

CHOLMOD_OK, CHOLMOD_NOT_INSTALLED, CHOLMOD_OUT_OF_MEMORY, CHOLMOD_TOO_LARGE, CHOLMOD_INVALID, CHOLMOD_GPU_PROBLEM = 0, -1, -2, -3, -4, -5
CHOLMOD_NOT_POSDEF, CHOLMOD_DSMALL = 1, 2
CHOLMOD_PATTERN, CHOLMOD_REAL, CHOLMOD_COMPLEX, CHOLMOD_ZOMPLEX = 0, 1, 2, 3
CHOLMOD_INT, CHOLMOD_LONG = 0, 2
CHOLMOD_SIMPLICIAL, CHOLMOD_AUTO, CHOLMOD_SUPERNODAL = 0, 1, 2
CHOLMOD_NATURAL, CHOLMOD_GIVEN, CHOLMOD_AMD = 0, 1, 2
CHOLMOD_A, CHOLMOD_LDLt, CHOLMOD_LD, CHOLMOD_DLt, CHOLMOD_L, CHOLMOD_Lt, CHOLMOD_D, CHOLMOD_P, CHOLMOD_Pt = range(9)


class Sparse(C.Structure):
    _fields_ = [("nrow", C.c_size_t), ("ncol", C.c_size_t), ("nzmax", C.c_size_t),
                ("p", C.c_void_p), ("i", C.c_void_p), ("nz", C.c_void_p), ("x", C.c_void_p), ("z", C.c_void_p),
                ("stype", C.c_int), ("itype", C.c_int), ("xtype", C.c_int), ("dtype", C.c_int),
                ("sorted", C.c_int), ("packed", C.c_int)]


class Dense(C.Structure):
    _fields_ = [("nrow", C.c_size_t), ("ncol", C.c_size_t), ("nzmax", C.c_size_t), ("d", C.c_size_t),
                ("x", C.c_void_p), ("z", C.c_void_p), ("xtype", C.c_int), ("dtype", C.c_int)]


class Factor(C.Structure):
    _fields_ = [("n", C.c_size_t), ("minor", C.c_size_t),
                ("Perm", C.c_void_p), ("ColCount", C.c_void_p), ("IPerm", C.c_void_p),
                ("nzmax", C.c_size_t),
                ("p", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p), ("z", C.c_void_p), ("nz", C.c_void_p),
                ("next", C.c_void_p), ("prev", C.c_void_p),
                ("nsuper", C.c_size_t), ("ssize", C.c_size_t), ("xsize", C.c_size_t),
                ("maxcsize", C.c_size_t), ("maxesize", C.c_size_t),
                ("super", C.c_void_p), ("pi", C.c_void_p), ("px", C.c_void_p), ("s", C.c_void_p),
                ("ordering", C.c_int), ("is_ll", C.c_int), ("is_super", C.c_int), ("is_monotonic", C.c_int),
                ("itype", C.c_int), ("xtype", C.c_int), ("dtype", C.c_int), ("useGPU", C.c_int)]


class Method(C.Structure):
    _fields_ = [("lnz", C.c_double), ("fl", C.c_double), ("prune_dense", C.c_double), ("prune_dense2", C.c_double),
                ("nd_oksep", C.c_double), ("other_1", C.c_double * 4),
                ("nd_small", C.c_size_t), ("other_2", C.c_size_t * 4),
                ("aggressive", C.c_int), ("order_for_lu", C.c_int), ("nd_compress", C.c_int), ("nd_camd", C.c_int),
                ("nd_components", C.c_int), ("ordering", C.c_int), ("other_3", C.c_size_t * 4)]


ERROR_HANDLER = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_int, C.c_char_p)


class Common(C.Structure):
    _fields_ = [("dbound", C.c_double), ("grow0", C.c_double), ("grow1", C.c_double),
                ("grow2", C.c_size_t), ("maxrank", C.c_size_t), ("supernodal_switch", C.c_double),
                ("supernodal", C.c_int), ("final_asis", C.c_int), ("final_super", C.c_int), ("final_ll", C.c_int),
                ("final_pack", C.c_int), ("final_monotonic", C.c_int), ("final_resymbol", C.c_int),
                ("zrelax", C.c_double * 3), ("nrelax", C.c_size_t * 3),
                ("prefer_zomplex", C.c_int), ("prefer_upper", C.c_int), ("quick_return_if_not_posdef", C.c_int),
                ("prefer_binary", C.c_int), ("print", C.c_int), ("precise", C.c_int), ("try_catch", C.c_int),
                ("error_handler", C.c_void_p),
                ("nmethods", C.c_int), ("current", C.c_int), ("selected", C.c_int),
                ("method", Method * (MAXMETHODS + 1)),
                ("postorder", C.c_int), ("default_nesdis", C.c_int),
                ("metis_memory", C.c_double), ("metis_dswitch", C.c_double), ("metis_nswitch", C.c_size_t),
                ("nrow", C.c_size_t), ("mark", c_long), ("iworksize", C.c_size_t), ("xworksize", C.c_size_t),
                ("Flag", C.c_void_p), ("Head", C.c_void_p), ("Xwork", C.c_void_p), ("Iwork", C.c_void_p),
                ("itype", C.c_int), ("dtype", C.c_int), ("no_workspace_reallocate", C.c_int), ("status", C.c_int),
                ("fl", C.c_double), ("lnz", C.c_double), ("anz", C.c_double), ("modfl", C.c_double),
                ("malloc_count", C.c_size_t), ("memory_usage", C.c_size_t), ("memory_inuse", C.c_size_t),
                ("nrealloc_col", C.c_double), ("nrealloc_factor", C.c_double), ("ndbounds_hit", C.c_double),
                ("rowfacfl", C.c_double), ("aatfl", C.c_double),
                ("called_nd", C.c_int), ("blas_ok", C.c_int),
                ("SPQR_grain", C.c_double), ("SPQR_small", C.c_double),
                ("SPQR_shrink", C.c_int), ("SPQR_nthreads", C.c_int),
                ("SPQR_flopcount", C.c_double), ("SPQR_analyze_time", C.c_double), ("SPQR_factorize_time", C.c_double),
                ("SPQR_solve_time", C.c_double), ("SPQR_flopcount_bound", C.c_double), ("SPQR_tol_used", C.c_double),
                ("SPQR_norm_E_fro", C.c_double), ("SPQR_istat", c_long * 10),
                ("useGPU", C.c_int), ("maxGpuMemBytes", C.c_size_t), ("maxGpuMemFraction", C.c_double),
                ("gpuMemorySize", C.c_size_t), ("gpuKernelTime", C.c_double), ("gpuFlops", c_long),
                ("gpuNumKernelLaunches", C.c_int),
                ("cublasHandle", C.c_void_p), ("gpuStream", C.c_void_p * 8), ("cublasEventPotrf", C.c_void_p * 3),
                ("updateCKernelsComplete", C.c_void_p), ("updateCBuffersFree", C.c_void_p * 8),
                ("dev_mempool", C.c_void_p), ("dev_mempool_size", C.c_size_t),
                ("host_pinned_mempool", C.c_void_p), ("host_pinned_mempool_size", C.c_size_t),
                ("devBuffSize", C.c_size_t), ("ibuffer", C.c_int), ("syrkStart", C.c_double),
                ("cpu_gemm_time", C.c_double), ("cpu_syrk_time", C.c_double), ("cpu_trsm_time", C.c_double),
                ("cpu_potrf_time", C.c_double), ("gpu_gemm_time", C.c_double), ("gpu_syrk_time", C.c_double),
                ("gpu_trsm_time", C.c_double), ("gpu_potrf_time", C.c_double),
                ("assemble_time", C.c_double), ("assemble_time2", C.c_double),
                ("cpu_gemm_calls", C.c_size_t), ("cpu_syrk_calls", C.c_size_t), ("cpu_trsm_calls", C.c_size_t),
                ("cpu_potrf_calls", C.c_size_t), ("gpu_gemm_calls", C.c_size_t), ("gpu_syrk_calls", C.c_size_t),
                ("gpu_trsm_calls", C.c_size_t), ("gpu_potrf_calls", C.c_size_t)]


assert C.sizeof(Common) == 2664 and C.sizeof(Factor) == 208 and C.sizeof(Sparse) == 88 and C.sizeof(Dense) == 56

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200_LIB = os.path.join(REPO, "suitesparse_b200", "csrc", "libsuitesparse_b200.so")


def default_host_cholmod() -> str:
    """The host libcholmod the application links: $SSB200_CHOLMOD_LIB, else this box's build of the unmodified
    reference (baseline/_ref/libcholmod.so, built by `make -C oracle host`; never anything under oracle/)."""
    return os.environ.get("SSB200_CHOLMOD_LIB", os.path.join(REPO, "baseline", "_ref", "libcholmod.so"))


_b200_handle = None


def load_b200(path: str | None = None) -> C.CDLL:
    """Load the B200 library globally so its three hot-path symbols interpose libcholmod's.  No fallback."""
    global _b200_handle
    if _b200_handle is None:
        path = path or B200_LIB
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback for the hot path")
        _b200_handle = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _b200_handle


def _np_view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class Cholmod:
    """One cholmod_common plus the handful of API calls of the demo program (Demo/cholmod_l_demo.c:98-331).

    gpu=True loads libsuitesparse_b200.so first (interposed hot path); gpu=False is the stock CPU library.
    Because interposition is decided when the host library binds its PLT, one process should construct all
    its Cholmod objects with the same `gpu` value *before* the first factorize, or use `ref_super_numeric`.
    """

    def __init__(self, gpu: bool, host_lib: str | None = None):
        self.gpu = gpu
        self.b200 = load_b200() if gpu else None
        host_lib = host_lib or default_host_cholmod()
        if not os.path.exists(host_lib):
            raise RuntimeError(f"host libcholmod not found at {host_lib} (set SSB200_CHOLMOD_LIB)")
        self.lib = C.CDLL(host_lib, mode=C.RTLD_GLOBAL)
        L = self.lib
        P = C.POINTER
        L.cholmod_l_start.argtypes = [P(Common)]
        L.cholmod_l_finish.argtypes = [P(Common)]
        L.cholmod_l_analyze.argtypes = [P(Sparse), P(Common)]; L.cholmod_l_analyze.restype = P(Factor)
        L.cholmod_l_analyze_p.argtypes = [P(Sparse), C.c_void_p, C.c_void_p, C.c_size_t, P(Common)]
        L.cholmod_l_analyze_p.restype = P(Factor)
        L.cholmod_l_factorize.argtypes = [P(Sparse), P(Factor), P(Common)]
        L.cholmod_l_factorize_p.argtypes = [P(Sparse), P(C.c_double), C.c_void_p, C.c_size_t, P(Factor), P(Common)]
        L.cholmod_l_solve.argtypes = [C.c_int, P(Factor), P(Dense), P(Common)]; L.cholmod_l_solve.restype = P(Dense)
        L.cholmod_l_solve2.argtypes = [C.c_int, P(Factor), P(Dense), C.c_void_p, P(P(Dense)), C.c_void_p,
                                       P(P(Dense)), P(P(Dense)), P(Common)]
        L.cholmod_l_free_factor.argtypes = [P(P(Factor)), P(Common)]
        L.cholmod_l_free_dense.argtypes = [P(P(Dense)), P(Common)]
        L.cholmod_l_free_sparse.argtypes = [P(P(Sparse)), P(Common)]
        L.cholmod_l_ptranspose.argtypes = [P(Sparse), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, P(Common)]
        L.cholmod_l_ptranspose.restype = P(Sparse)
        L.cholmod_l_transpose.argtypes = [P(Sparse), C.c_int, P(Common)]; L.cholmod_l_transpose.restype = P(Sparse)
        L.cholmod_l_check_factor.argtypes = [P(Factor), P(Common)]
        L.cholmod_l_copy_factor.argtypes = [P(Factor), P(Common)]; L.cholmod_l_copy_factor.restype = P(Factor)
        L.cholmod_l_change_factor.argtypes = [C.c_int] * 5 + [P(Factor), P(Common)]
        L.cholmod_l_rcond.argtypes = [P(Factor), P(Common)]; L.cholmod_l_rcond.restype = C.c_double
        L.cholmod_l_allocate_dense.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, P(Common)]
        L.cholmod_l_allocate_dense.restype = P(Dense)
        for name in ("cholmod_l_super_numeric",):
            getattr(L, name).argtypes = [P(Sparse), P(Sparse), P(C.c_double), P(Factor), P(Common)]
        for name in ("cholmod_l_super_lsolve", "cholmod_l_super_ltsolve"):
            getattr(L, name).argtypes = [P(Factor), P(Dense), P(Dense), P(Common)]
        self.cm = Common()
        L.cholmod_l_start(C.byref(self.cm))
        self.cm.print = 0                       # errors are reported through status; keep stdout clean
        self._keep = []

    # -- the three hot-path entry points as THIS process resolves them (interposed or stock) -----------------
    def hot(self, name: str):
        if self.gpu:
            f = getattr(self.b200, name)
            f.argtypes = getattr(self.lib, name).argtypes
            return f
        return getattr(self.lib, name)

    def finish(self):
        self.lib.cholmod_l_finish(C.byref(self.cm))

    # -- objects ---------------------------------------------------------------------------------------------
    def sparse(self, A, stype: int, zomplex: bool = False) -> Sparse:
        """Wrap a scipy CSC matrix (no copy).  stype=+1: upper triangle stored, -1: lower, 0: unsymmetric.
        A complex matrix becomes CHOLMOD_COMPLEX (interleaved re,im) or, with zomplex=True, CHOLMOD_ZOMPLEX (x and z arrays)."""
        Ap = np.ascontiguousarray(A.indptr, dtype=np.int64)
        Ai = np.ascontiguousarray(A.indices, dtype=np.int64)
        S = Sparse()
        S.nz = None; S.z = None
        xtype = CHOLMOD_REAL
        if np.iscomplexobj(A.data):
            data = np.ascontiguousarray(A.data, dtype=np.complex128)
            if zomplex:
                Ax = np.ascontiguousarray(data.real); Az = np.ascontiguousarray(data.imag)
                self._keep.append(Az); S.z = Az.ctypes.data; xtype = CHOLMOD_ZOMPLEX
            else:
                Ax = data.view(np.float64); xtype = CHOLMOD_COMPLEX
        else:
            Ax = np.ascontiguousarray(A.data, dtype=np.float64)
        self._keep.append((Ap, Ai, Ax))
        S.nrow, S.ncol, S.nzmax = A.shape[0], A.shape[1], max(1, Ai.size)
        S.p, S.i, S.x = Ap.ctypes.data, Ai.ctypes.data, Ax.ctypes.data
        S.stype, S.itype, S.xtype, S.dtype, S.sorted, S.packed = stype, CHOLMOD_LONG, xtype, 0, 1, 1
        return S

    def dense(self, X: np.ndarray) -> Dense:
        assert X.dtype in (np.float64, np.complex128) and (X.flags.f_contiguous or X.ndim == 1)
        X2 = X.reshape(X.shape[0], -1, order="F")
        self._keep.append(X2)
        D = Dense()
        D.nrow, D.ncol, D.d = X2.shape[0], X2.shape[1], X2.shape[0]
        D.nzmax = max(1, X2.size)
        D.x, D.z, D.xtype, D.dtype = X2.ctypes.data, None, (CHOLMOD_COMPLEX if X.dtype == np.complex128 else CHOLMOD_REAL), 0
        return D

    # -- driver calls ----------------------------------------------------------------------------------------
    def analyze(self, S: Sparse, perm: np.ndarray | None = None, supernodal: int = CHOLMOD_SUPERNODAL,
                nrelax=None, zrelax=None, postorder: bool = True):
        cm = self.cm
        cm.supernodal = supernodal
        cm.postorder = 1 if postorder else 0
        if nrelax is not None:
            for t in range(3):
                cm.nrelax[t] = nrelax[t]
        if zrelax is not None:
            for t in range(3):
                cm.zrelax[t] = zrelax[t]
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int64)
            self._keep.append(perm)
            cm.nmethods = 1
            cm.method[0].ordering = CHOLMOD_GIVEN
            Lp = self.lib.cholmod_l_analyze_p(C.byref(S), perm.ctypes.data, None, 0, C.byref(cm))
        else:
            Lp = self.lib.cholmod_l_analyze(C.byref(S), C.byref(cm))
        if not Lp:
            raise RuntimeError(f"cholmod_l_analyze failed, status {cm.status}")
        return Lp

    def factorize(self, S: Sparse, Lp, beta: float | None = None) -> int:
        if beta is None:
            ok = self.lib.cholmod_l_factorize(C.byref(S), Lp, C.byref(self.cm))
        else:
            b = (C.c_double * 2)(beta, 0.0)
            ok = self.lib.cholmod_l_factorize_p(C.byref(S), b, None, 0, Lp, C.byref(self.cm))
        return ok

    def solve(self, Lp, B: np.ndarray, system: int = CHOLMOD_A) -> np.ndarray:
        cplx = np.iscomplexobj(B)
        Bd = self.dense(np.asfortranarray(B, dtype=np.complex128 if cplx else np.float64))
        # a C program resolves cholmod_l_solve to the interposed definition (device-side P, L, L', P' for the main case, else
        # the host library's); ctypes resolves per handle, so pick the same one explicitly
        lib = self.b200 if self.gpu else self.lib
        if self.gpu:
            lib.cholmod_l_solve.argtypes = self.lib.cholmod_l_solve.argtypes; lib.cholmod_l_solve.restype = self.lib.cholmod_l_solve.restype
        Xp = lib.cholmod_l_solve(system, Lp, C.byref(Bd), C.byref(self.cm))
        if not Xp:
            raise RuntimeError(f"cholmod_l_solve failed, status {self.cm.status}")
        dt = np.complex128 if Xp.contents.xtype == CHOLMOD_COMPLEX else np.float64
        X = _np_view(Xp.contents.x, Xp.contents.nzmax, dt)[: Xp.contents.nrow * Xp.contents.ncol].copy()
        X = X.reshape((Xp.contents.nrow, Xp.contents.ncol), order="F")
        pp = C.POINTER(Dense)(Xp.contents)
        self.lib.cholmod_l_free_dense(C.byref(pp), C.byref(self.cm))
        return X if B.ndim > 1 else X[:, 0]

    def lower_permuted(self, S: Sparse, Lp):
        """S2 = tril(P A P') exactly as cholmod_factorize_p builds it for a symmetric-upper A
        (Cholesky/cholmod_factorize.c:222-231): one ptranspose with values and L->Perm."""
        assert S.stype > 0
        return self.lib.cholmod_l_ptranspose(C.byref(S), 2, Lp.contents.Perm, None, 0, C.byref(self.cm))

    def free_factor(self, Lp):
        pp = C.POINTER(Factor)(Lp.contents)
        # a C program resolves cholmod_l_free_factor to the interposed definition (which drops the cached plan and then
        # calls the host library's); ctypes resolves per handle, so pick the same one explicitly
        lib = self.b200 if self.gpu else self.lib
        lib.cholmod_l_free_factor.argtypes = [C.POINTER(C.POINTER(Factor)), C.POINTER(Common)]
        lib.cholmod_l_free_factor(C.byref(pp), C.byref(self.cm))

    def free_sparse(self, Sp):
        pp = C.POINTER(Sparse)(Sp.contents)
        self.lib.cholmod_l_free_sparse(C.byref(pp), C.byref(self.cm))

    # -- views of the factor -----------------------------------------------------------------------------------
    @staticmethod
    def factor_arrays(Lp) -> dict:
        L = Lp.contents
        ns = L.nsuper
        out = dict(n=L.n, minor=L.minor, nsuper=ns, ssize=L.ssize, xsize=L.xsize, maxcsize=L.maxcsize,
                   maxesize=L.maxesize, is_super=L.is_super, is_ll=L.is_ll, xtype=L.xtype, ordering=L.ordering,
                   Perm=_np_view(L.Perm, L.n, np.int64), ColCount=_np_view(L.ColCount, L.n, np.int64))
        if L.is_super:
            out["super"] = _np_view(L.super, ns + 1, np.int64)
            out["pi"] = _np_view(L.pi, ns + 1, np.int64)
            out["px"] = _np_view(L.px, ns + 1, np.int64)
            out["s"] = _np_view(L.s, int(out["pi"][ns]) if ns else 0, np.int64)
            out["x"] = (_np_view(L.x, L.xsize, np.complex128 if L.xtype == CHOLMOD_COMPLEX else np.float64)
                        if L.xtype != CHOLMOD_PATTERN else None)
        return out
