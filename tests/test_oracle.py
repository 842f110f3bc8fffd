"""CPU tests pinning the oracle (oracle/ssb_oracle.c): against the golden fixtures generated from the reference build
(tests/golden/make_golden.py), against the reference's pinned symbolic numbers for mesh problems
(MATLAB_Tools/MESHND/meshnd_quality_out.txt) and, when present on this box, against the reference library itself."""
import os
import numpy as np
import pytest
import scipy.sparse as sp
from conftest import GOLDEN, load_golden, golden_matrix, persuper_relerr, REF_LIB
from oracle import oracle

TOL_L = 1e-11      # max |L - L_ref| / max|L_ref| per supernode; BLAS-dependent rounding only


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g = load_golden(path)
    S, F = golden_matrix(g)
    st, minor, Lx = oracle.factorize(int(g["n"]), g["super"], g["pi"], g["px"], g["s"], S, beta=float(g["beta"]),
                                     quick_return=bool(g.get("quick", 0)), F=F)
    assert st == int(g["status"])
    assert minor == int(g["minor"])
    assert persuper_relerr(g["px"], Lx, g["Lx"]) < TOL_L
    # the strictly upper part of every diagonal block stays exactly zero
    for s in range(len(g["px"]) - 1):
        nscol = int(g["super"][s + 1] - g["super"][s]); nsrow = int(g["pi"][s + 1] - g["pi"][s])
        blk = Lx[int(g["px"][s]):int(g["px"][s]) + nsrow * nscol].reshape((nsrow, nscol), order="F")
        assert np.all(np.triu(blk[:nscol, :], 1) == 0.0)
    up = oracle.enumerate_updates(int(g["n"]), g["super"], g["pi"], g["s"])
    if int(g["status"]) == 0 and int(g["syrk_calls"]) > 0:
        assert len(up["d"]) == int(g["syrk_calls"])          # one dsyrk per (d,s) update in the reference
    assert up["maxcsize"] == int(g["maxcsize"])
    if g["x"].size:
        perm = g["Perm"]
        y = oracle.lsolve(g["super"], g["pi"], g["px"], g["s"], Lx, g["b"][perm])
        y = oracle.lsolve(g["super"], g["pi"], g["px"], g["s"], Lx, y, transpose=True)
        x = np.empty_like(y); x[perm] = y
        assert np.abs(x - g["x"]).max() <= 1e-9 * max(1.0, np.abs(g["x"]).max())


def test_mesh_symbolic_numbers_pinned_by_reference():
    """meshnd_quality_out.txt:691-694: 3D 7-point 64^3 with MESHND: nnz(L)=1.566e8, flops=4.141e11; smaller meshes here."""
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build not present")
    from suitesparse_b200 import gen
    from suitesparse_b200.cholmod_host import Cholmod
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem("lap7", 32)
    L = ch.analyze(ch.sparse(A, +1), p)
    # 32^3 is not in the table; the 64^3 entry is checked in the GPU suite's full-size run.  Here: internal consistency
    f = ch.factor_arrays(L)
    assert abs(ch.cm.lnz - float((f["ColCount"]).sum())) < 1
    assert abs(ch.cm.fl - float((f["ColCount"].astype(np.float64) ** 2).sum())) < 1e-6 * ch.cm.fl
    ch.free_factor(L)


@pytest.mark.parametrize("kind,N", [("lap7", 10), ("lap27", 8), ("elas", 5)])
def test_oracle_matches_reference_library_live(kind, N):
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build not present")
    from suitesparse_b200 import gen
    from suitesparse_b200.cholmod_host import Cholmod, _np_view
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem(kind, N)
    S = ch.sparse(A, +1)
    L = ch.analyze(S, p)
    assert ch.factorize(S, L) and ch.cm.status == 0
    f = ch.factor_arrays(L)
    S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
    Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
    st, minor, Lx = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], sp.csc_matrix((Ax, Ai, Ap), shape=(n, n)))
    assert st == 0 and minor == n
    assert persuper_relerr(f["px"], Lx, f["x"]) < TOL_L
    b = np.ones(n)
    x = ch.solve(L, b)
    y = oracle.lsolve(f["super"], f["pi"], f["px"], f["s"], Lx, b[f["Perm"]])
    y = oracle.lsolve(f["super"], f["pi"], f["px"], f["s"], Lx, y, transpose=True)
    xo = np.empty(n); xo[f["Perm"]] = y
    assert np.abs(x - xo).max() < 1e-10 * np.abs(x).max()
    ch.free_sparse(S2); ch.free_factor(L)
