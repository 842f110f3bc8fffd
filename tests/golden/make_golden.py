"""Generate the golden fixtures under tests/golden/ from the REFERENCE build (oracle/_ref/libcholmod_ref.so, compiled
from /root/reference by oracle/Makefile).  Run here (needs /root/reference for the bundled matrices):

    PYTHONPATH=. python tests/golden/make_golden.py

For every bundled real matrix it stores the exact arguments cholmod_l_factorize hands to cholmod_l_super_numeric
(S = tril(P A P') for symmetric input; S = A(p,:), F = S' and beta for unsymmetric input, Cholesky/cholmod_factorize.c:186-256)
and what the reference's CPU+BLAS path returns: super/pi/px/s, L->x, L->minor, Common->status, and the solution of
A x = b (or (AA'+beta I) x = b) for the demo's right-hand side b_i = 1 + i/n (Demo/cholmod_l_demo.c:231-239).
The mesh problems pin nnz(L)/flops against MATLAB_Tools/MESHND/meshnd_quality_out.txt (tests/test_oracle.py).
"""
import ctypes as C, os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from suitesparse_b200.cholmod_host import Cholmod, Sparse, _np_view, CHOLMOD_SUPERNODAL, CHOLMOD_REAL

REF = os.environ.get("SSB200_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]; libc.fclose.argtypes = [C.c_void_p]

CASES = [  # (file, relax) relax: "default" or "none" (nrelax=zrelax=0, Tcov/solve.c:105-120)
    ("CHOLMOD/Demo/Matrix/bcsstk01.tri", "default"), ("CHOLMOD/Demo/Matrix/bcsstk01.tri", "none"),
    ("CHOLMOD/Demo/Matrix/bcsstk02.tri", "default"), ("CHOLMOD/Demo/Matrix/can___24.mtx", "default"),
    ("CHOLMOD/Demo/Matrix/pts5ldd03.mtx", "default"), ("CHOLMOD/Demo/Matrix/pts5ldd03.mtx", "none"),
    ("CHOLMOD/Demo/Matrix/lp_afiro.tri", "default"), ("CHOLMOD/Demo/Matrix/one.tri", "default"), ("CHOLMOD/Demo/Matrix/two.tri", "default"),
    ("CHOLMOD/Tcov/Matrix/k01up", "default"), ("CHOLMOD/Tcov/Matrix/ex5lo", "default"), ("CHOLMOD/Tcov/Matrix/20lo", "default"),
    ("CHOLMOD/Tcov/Matrix/r5lo", "default"), ("CHOLMOD/Tcov/Matrix/4lo", "default"), ("CHOLMOD/Tcov/Matrix/ibm32", "default"),
    ("CHOLMOD/Tcov/Matrix/galenet", "default"), ("CHOLMOD/Tcov/Matrix/5by50", "default"), ("CHOLMOD/Tcov/Matrix/3singular", "default"),
    ("CHOLMOD/Tcov/Matrix/2diag.tri", "default"), ("CHOLMOD/Tcov/Matrix/plskz362.mtx", "default"),
]


def to_scipy(Sp):
    s = Sp.contents
    Ap = _np_view(s.p, s.ncol + 1, np.int64).copy()
    if s.packed:
        nz = int(Ap[s.ncol]); Ai = _np_view(s.i, nz, np.int64).copy(); Ax = _np_view(s.x, nz, np.float64).copy()
        M = sp.csc_matrix((Ax, Ai, Ap), shape=(s.nrow, s.ncol))
    else:
        raise RuntimeError("unpacked")
    return M, s.stype, s.xtype


def main():
    made = []
    for rel, relax in CASES:
        path = os.path.join(REF, rel)
        ch = Cholmod(gpu=False)
        ch.lib.cholmod_l_read_sparse.restype = C.POINTER(Sparse); ch.lib.cholmod_l_read_sparse.argtypes = [C.c_void_p, C.c_void_p]
        fp = libc.fopen(path.encode(), b"r")
        Aptr = ch.lib.cholmod_l_read_sparse(fp, C.byref(ch.cm)); libc.fclose(fp)
        if not Aptr:
            print("skip (unreadable)", rel); continue
        A, stype, xtype = to_scipy(Aptr)
        if xtype != CHOLMOD_REAL:
            print("skip (not real)", rel); continue
        n = A.shape[0]
        ch.cm.supernodal = CHOLMOD_SUPERNODAL
        if relax == "none":
            for t in range(3): ch.cm.nrelax[t] = 0; ch.cm.zrelax[t] = 0.0
        Lp = ch.lib.cholmod_l_analyze(Aptr, C.byref(ch.cm))
        if not Lp or not Lp.contents.is_super:
            print("skip (analyze)", rel, ch.cm.status); continue
        beta = 0.0 if stype != 0 else 1e-6           # Demo/cholmod_l_demo.c:279-293
        b2 = (C.c_double * 2)(beta, 0.0)
        ok = ch.lib.cholmod_l_factorize_p(Aptr, b2, None, 0, Lp, C.byref(ch.cm))
        status = ch.cm.status
        f = ch.factor_arrays(Lp)
        perm = f["Perm"].copy()
        # arguments of super_numeric, rebuilt in scipy exactly as cholmod_factorize_p does
        if stype != 0:
            full = A + (sp.tril(A, -1).T if stype < 0 else sp.triu(A, 1).T)
            Pm = full.tocsc()[perm, :][:, perm]
            S = sp.tril(Pm).tocsc(); S.sort_indices(); F = None
        else:
            S = A.tocsc()[perm, :].tocsc(); S.sort_indices()
            F = S.T.tocsc(); F.sort_indices()
        bvec = 1.0 + np.arange(n) / n
        x = None
        if status == 0:
            x = ch.solve(Lp, bvec)
        name = os.path.basename(rel).replace(".", "_") + ("_norelax" if relax == "none" else "")
        d = dict(name=name, source=rel, relax=relax, n=n, stype=(-1 if stype != 0 else 0), beta=beta, ok=ok, status=status, minor=f["minor"],
                 Sp=S.indptr.astype(np.int64), Si=S.indices.astype(np.int64), Sx=S.data, ncolS=S.shape[1],
                 Perm=perm, super=f["super"].copy(), pi=f["pi"].copy(), px=f["px"].copy(), s=f["s"].copy(), Lx=f["x"].copy(),
                 maxcsize=f["maxcsize"], maxesize=f["maxesize"], b=bvec, x=(x if x is not None else np.zeros(0)),
                 syrk_calls=ch.cm.cpu_syrk_calls, potrf_calls=ch.cm.cpu_potrf_calls)
        if F is not None:
            d.update(Fp=F.indptr.astype(np.int64), Fi=F.indices.astype(np.int64), Fx=F.data)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        made.append((name, n, int(f["nsuper"]), int(f["xsize"]), status, int(f["minor"]), int(ch.cm.cpu_syrk_calls)))
        ch.free_factor(Lp)
    # not-positive-definite protocol (t_cholmod_super_numeric.c:905-968): a mesh Laplacian with one diagonal entry made
    # negative, so that the reference stops in the middle of a supernode (minor != first column) or at its first column.
    from suitesparse_b200 import gen
    A0, perm0 = gen.make_problem("lap7", 6)
    for tag, (kbad, val, quick) in {"npd_mid": (150, -5.0, 0), "npd_mid_quick": (150, -5.0, 1), "npd_last": (215, -50.0, 0),
                                    "npd_first": (0, -1.0, 0), "npd_zero_pivot": (77, 0.0, 0)}.items():
        ch = Cholmod(gpu=False)
        A = A0.copy().tolil(); A[kbad, kbad] = val; A = A.tocsc(); A.sort_indices()
        A.indices = A.indices.astype(np.int64); A.indptr = A.indptr.astype(np.int64)
        S_up = ch.sparse(A, +1)
        ch.cm.quick_return_if_not_posdef = quick
        Lp = ch.analyze(S_up, perm0)
        ok = ch.factorize(S_up, Lp)
        f = ch.factor_arrays(Lp)
        perm = f["Perm"].copy()
        full = A + sp.triu(A, 1).T
        S = sp.tril(full.tocsc()[perm, :][:, perm]).tocsc(); S.sort_indices()
        n = A.shape[0]
        d = dict(name=tag, source="gen.laplacian(6,7) with A[%d,%d]=%g" % (kbad, kbad, val), relax="default", n=n, stype=-1, beta=0.0, ok=ok,
                 status=ch.cm.status, minor=f["minor"], quick=quick, Sp=S.indptr.astype(np.int64), Si=S.indices.astype(np.int64), Sx=S.data, ncolS=n,
                 Perm=perm, super=f["super"].copy(), pi=f["pi"].copy(), px=f["px"].copy(), s=f["s"].copy(), Lx=f["x"].copy(),
                 maxcsize=f["maxcsize"], maxesize=f["maxesize"], b=np.zeros(0), x=np.zeros(0), syrk_calls=0, potrf_calls=0)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **d)
        made.append((tag, n, int(f["nsuper"]), int(f["xsize"]), ch.cm.status, int(f["minor"]), 0))
        ch.free_factor(Lp)
    for m in made:
        print("%-28s n=%-5d nsuper=%-4d xsize=%-7d status=%d minor=%d syrk=%d" % m)


if __name__ == "__main__":
    main()
