"""Complex / zomplex fixtures under tests/golden_complex/ from the REFERENCE build (its zherk/zgemm/zpotrf/ztrsm instantiations,
CHOLMOD/Supernodal/cholmod_super_numeric.c:81-86, t_cholmod_super_numeric.c:41-83):

    PYTHONPATH=. python tests/golden/make_golden_complex.py

Hermitian positive definite test matrices: a mesh Laplacian (diagonal raised to 8) plus i*0.3*S with S skew-symmetric on the
mesh edges; an unsymmetric complex A for the A*A'+beta*I path; a case that is not positive definite.  Stored: the matrix as the
application passes it (upper triangle / unsymmetric, complex CSC), the permutation, and what the reference returns."""
import ctypes as C, os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from suitesparse_b200 import gen
from suitesparse_b200.cholmod_host import Cholmod, CHOLMOD_SUPERNODAL

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "golden_complex")


def hermitian_mesh(kind, N, shift=2.0, imag=0.3):
    A, perm = gen.make_problem(kind, N)                       # upper triangle, real
    A = A.tocoo()
    off = A.row != A.col
    rng = np.random.default_rng(42)
    sign = np.where(rng.random(A.nnz) < 0.5, -1.0, 1.0)
    data = A.data.astype(np.complex128)
    data[off] += 1j * imag * sign[off]                        # upper entries a + ib; the lower ones are their conjugates
    data[~off] += shift
    U = sp.csc_matrix((data, (A.row, A.col)), shape=A.shape); U.sort_indices()
    U.indices = U.indices.astype(np.int64); U.indptr = U.indptr.astype(np.int64)
    return U, perm


def run(name, U, perm, stype, beta=None, zomplex=False, bad=None, b=None):
    ch = Cholmod(gpu=False)
    if bad is not None:
        U = U.tolil(); U[bad, bad] = -3.0; U = U.tocsc(); U.sort_indices()
        U.indices = U.indices.astype(np.int64); U.indptr = U.indptr.astype(np.int64)
    S = ch.sparse(U, stype, zomplex=zomplex)
    ch.cm.supernodal = CHOLMOD_SUPERNODAL
    if perm is not None:
        Lp = ch.analyze(S, perm)
    else:
        Lp = ch.lib.cholmod_l_analyze(C.byref(S), C.byref(ch.cm))
    ok = ch.factorize(S, Lp, beta=beta)
    status = ch.cm.status
    f = ch.factor_arrays(Lp)
    n = U.shape[0]
    x = np.zeros(0, dtype=np.complex128)
    if b is None:
        b = (1.0 + np.arange(n) / n) + 1j * (0.5 - np.arange(n) / (2.0 * n))
    if status == 0:
        x = ch.solve(Lp, b)
    d = dict(name=name, n=n, stype=stype, beta=(0.0 if beta is None else beta), zomplex=int(zomplex), ok=ok, status=status, minor=f["minor"],
             Up=U.indptr.astype(np.int64), Ui=U.indices.astype(np.int64), Ux=U.data.astype(np.complex128), ncolU=U.shape[1],
             perm=(perm if perm is not None else np.zeros(0, dtype=np.int64)), Perm=f["Perm"].copy(),
             super=f["super"].copy(), pi=f["pi"].copy(), px=f["px"].copy(), s=f["s"].copy(), Lx=f["x"].copy(),
             maxcsize=f["maxcsize"], maxesize=f["maxesize"], b=b, x=x, potrf_calls=ch.cm.cpu_potrf_calls)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-24s n=%-5d nsuper=%-4d xsize=%-7d status=%d minor=%d" % (name, n, f["nsuper"], f["xsize"], status, f["minor"]))
    ch.free_factor(Lp)


def main():
    U, perm = hermitian_mesh("lap7", 7)
    run("herm_lap7_7", U, perm, +1)
    run("herm_lap7_7_zomplex", U, perm, +1, zomplex=True)
    U2, perm2 = hermitian_mesh("lap27", 6, shift=4.0)
    run("herm_lap27_6", U2, perm2, +1)
    U3, perm3 = hermitian_mesh("lap7", 12)
    run("herm_lap7_12", U3, perm3, +1)                         # supernodes wider than 64 columns, several levels
    run("herm_lap7_7_npd", U, perm, +1, bad=170)
    rng = np.random.default_rng(42)
    m, k = 40, 60
    R = sp.random(m, k, density=0.1, random_state=rng, format="csc")
    I = sp.random(m, k, density=0.1, random_state=rng, format="csc")
    A = (R + 1j * I + sp.eye(m, k, format="csc")).tocsc(); A.sort_indices()
    A.indices = A.indices.astype(np.int64); A.indptr = A.indptr.astype(np.int64)
    run("unsym_40x60", A, None, 0, beta=1e-3)


if __name__ == "__main__":
    main()
