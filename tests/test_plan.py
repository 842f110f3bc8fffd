"""CPU tests of the host plan builder (ssb_plan.cpp): the precomputed update lists, levels and schedules are checked
against the oracle's enumeration and against first principles on the golden structures."""
import ctypes as C, os
import numpy as np
import pytest
from conftest import GOLDEN, load_golden, B200_LIB
from oracle import oracle


def plan_summary(n, super_, pi, px, s):
    lib = C.CDLL(B200_LIB)
    out = np.zeros(20); lev = np.zeros(len(super_) - 1, dtype=np.int32)
    a = [np.ascontiguousarray(v, dtype=np.int64) for v in (super_, pi, px, s)]
    rc = lib.ssb200_plan_summary(C.c_int64(n), C.c_int64(len(super_) - 1), *[v.ctypes.data_as(C.c_void_p) for v in a],
                                 out.ctypes.data_as(C.c_void_p), lev.ctypes.data_as(C.c_void_p))
    return rc, out, lev


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_plan_matches_oracle_enumeration(path):
    g = load_golden(path)
    n = int(g["n"])
    rc, out, lev = plan_summary(n, g["super"], g["pi"], g["px"], g["s"])
    assert rc == 0
    up = oracle.enumerate_updates(n, g["super"], g["pi"], g["s"])
    assert int(out[1]) == len(up["d"])
    assert int(out[2]) == int(up["ndrow2"].sum())
    # levels: every update's descendant lies strictly below its target; some supernode sits at every level
    assert np.all(lev[up["d"]] < lev[up["s"]])
    assert set(lev.tolist()) == set(range(int(out[0])))
    # dense flop counts from first principles
    nscol = np.diff(g["super"]).astype(np.float64); nsrow = np.diff(g["pi"]).astype(np.float64)
    assert np.isclose(out[12], (nscol ** 3 / 3).sum())
    assert np.isclose(out[13], (nscol ** 2 * (nsrow - nscol)).sum())
    ndcol = nscol[up["d"]]
    tri = up["ndrow1"] * up["ndrow2"] - 0.5 * up["ndrow1"] * (up["ndrow1"] - 1)
    assert np.isclose(out[11], (2 * ndcol * tri).sum())
    # the host-streaming copy tasks cover every entry of Lx exactly once
    assert int(out[16]) == int(g["px"][-1])
    # one potrf job per 64-column block, one solve job per block
    blocks = np.ceil(nscol / 64).sum()
    assert int(out[6]) == int(blocks) and int(out[10]) == int(blocks)


def test_plan_rejects_bad_structure():
    g = load_golden([p for p in GOLDEN if "bcsstk01_tri.npz" in p][0])
    s_bad = g["s"].copy(); s_bad[-1] = s_bad[-2]           # unsorted / duplicate row index
    rc, _, _ = plan_summary(int(g["n"]), g["super"], g["pi"], g["px"], s_bad)
    assert rc == -4
    sup_bad = g["super"].copy(); sup_bad[-1] += 1
    rc, _, _ = plan_summary(int(g["n"]), sup_bad, g["pi"], g["px"], g["s"])
    assert rc == -4


def test_generators():
    from suitesparse_b200 import gen
    A = gen.laplacian(5, 7)
    assert A.shape == (125, 125) and A.nnz == 125 + 3 * 5 * 5 * 4       # diagonal + one upper entry per mesh edge
    p = gen.meshnd_perm(5, 5, 5)
    assert sorted(p.tolist()) == list(range(125))
    # the last 25 ordered nodes are the middle plane k=2 (separator last, meshnd.m:96-100)
    assert set(p[-25:].tolist()) == set((np.arange(25) + 25 * 2).tolist())
    A27 = gen.laplacian(4, 27)
    full = A27 + __import__("scipy.sparse").sparse.triu(A27, 1).T
    assert np.allclose(full.diagonal(), 26.0) and abs(full.sum(axis=1).min()) >= 0
    E = gen.elasticity(4)
    assert E.shape == (192, 192)
    fullE = (E + __import__("scipy.sparse").sparse.triu(E, 1).T).toarray()
    assert np.linalg.eigvalsh(fullE).min() > 0


def test_three_level_blocking_and_copy_cover_on_a_wide_supernode():
    """A mesh whose root supernode is wider than 1024 columns: the in-supernode trailing updates use K = 64, 256 and 1024,
    the copy tasks still cover Lx exactly once, and the wide blocks get inverse slots."""
    import os
    from conftest import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build (host libcholmod for cholmod_l_analyze) not present")
    from suitesparse_b200 import gen
    from suitesparse_b200.cholmod_host import Cholmod
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem("lap7", 34)
    L = ch.analyze(ch.sparse(A, +1), p)
    f = ch.factor_arrays(L)
    assert np.diff(f["super"]).max() > 1024
    rc, out, lev = plan_summary(f["n"], f["super"], f["pi"], f["px"], f["s"])
    assert rc == 0
    assert int(out[19]) == 1024
    assert int(out[16]) == int(f["px"][-1])
    nscol = np.diff(f["super"]); nsrow = np.diff(f["pi"])
    wide = (nscol >= 33) & (nsrow > 32)
    assert int(out[18]) == int(np.ceil(nscol[wide] / 64).sum())
    ch.free_factor(L)


@pytest.mark.parametrize("prefer", [0, 1])
def test_lookahead_schedule_two_streams(prefer, monkeypatch):
    """The single-GPU look-ahead schedule (outer panel O+1 factorized on the panel stream while the main stream applies
    panel O to the rest): replayed on two emulated in-order streams with events, in both extreme interleavings, it must
    give the oracle's factor; in list order (look-ahead off) as well."""
    import scipy.sparse as sp
    from conftest import REF_LIB, persuper_relerr
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build (host libcholmod for cholmod_l_analyze) not present")
    import emulate_plan as E
    from suitesparse_b200 import gen
    from suitesparse_b200.cholmod_host import Cholmod, _np_view
    monkeypatch.setenv("SSB200_NB_OUTER", "256")          # three or more outer panels on a 24^3 mesh (root > 576 columns)
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem("lap7", 24)
    S = ch.sparse(A, +1); L = ch.analyze(S, p)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ch.factor_arrays(L).items()}
    assert np.diff(f["super"]).max() > 3 * 256
    S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
    Ap = _np_view(s2.p, n + 1, np.int64).copy(); Ai = _np_view(s2.i, int(Ap[n]), np.int64).copy(); Ax = _np_view(s2.x, int(Ap[n]), np.float64).copy()
    Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    st, minor, Lo = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], Sl)
    plan = E.export_plan(n, f["super"], f["pi"], f["px"], f["s"])
    Ln = plan["launches"]
    assert (Ln[:, 4] == 1).any() and (Ln[:, 0] == 5).any()           # a panel stream and its joins exist
    # every event is recorded exactly once, before (in list order) the launch that waits for it
    rec = {int(e): t for t, e in enumerate(Ln[:, 6]) if e >= 0}
    assert len(rec) == (Ln[:, 6] >= 0).sum()
    for t, e in enumerate(Ln[:, 5]):
        if e >= 0:
            assert rec[int(e)] < t
    rel = E.relmap_of(plan, f["pi"], f["s"])
    Lx = np.zeros(int(f["px"][-1]))
    E.assemble(plan, f["super"], f["pi"], f["px"], f["s"], Sl, Lx)
    reordered = E.run_two_streams(plan, rel, Lx, prefer)
    assert persuper_relerr(f["px"], Lx, Lo) < 1e-11
    if prefer == 1:
        assert reordered > 0                                          # the panel chain really ran ahead of main-stream work
    Lx2 = np.zeros(int(f["px"][-1]))
    E.assemble(plan, f["super"], f["pi"], f["px"], f["s"], Sl, Lx2)
    E.run_launches(plan, rel, Lx2, 0, len(Ln))
    assert persuper_relerr(f["px"], Lx2, Lo) < 1e-11
    ch.free_sparse(S2); ch.free_factor(L)
