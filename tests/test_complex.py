"""Complex / zomplex supernodal factorization (reference: the zherk/zgemm/zpotrf/ztrsm instantiations,
cholmod_super_numeric.c:81-86).  The CUDA path factorizes the real matrix of order 2n in which every complex entry a+ib is the
2x2 block [a -b; b a]: CPU tests pin the fixtures generated from the reference and the identity the path rests on; the GPU tests
compare the drop-in symbols with the reference's complex factor."""
import ctypes as C, glob, os
import numpy as np
import pytest
import scipy.sparse as sp
from conftest import REPO, B200_LIB, REF_LIB

CG = sorted(glob.glob(os.path.join(REPO, "tests", "golden_complex", "*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    return {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}


def dense_L(g):
    """The complex supernodal factor as a dense lower-triangular matrix (in the permuted ordering)."""
    n = int(g["n"]); L = np.zeros((n, n), dtype=np.complex128)
    for s in range(len(g["super"]) - 1):
        k1, k2 = int(g["super"][s]), int(g["super"][s + 1])
        rows = g["s"][int(g["pi"][s]):int(g["pi"][s + 1])]
        blk = g["Lx"][int(g["px"][s]):int(g["px"][s]) + len(rows) * (k2 - k1)].reshape((len(rows), k2 - k1), order="F")
        L[rows, k1:k2] = blk
    return np.tril(L)


def full_matrix(g):
    U = sp.csc_matrix((g["Ux"], g["Ui"], g["Up"]), shape=(int(g["n"]), int(g["ncolU"])))
    if g["stype"] != 0:
        return (U + sp.triu(U, 1).conj().T).toarray()
    A = U.toarray()
    return A @ A.conj().T + float(g["beta"]) * np.eye(A.shape[0])


def blockify(M):
    n, m = M.shape
    B = np.zeros((2 * n, 2 * m))
    B[0::2, 0::2] = M.real; B[1::2, 1::2] = M.real; B[1::2, 0::2] = M.imag; B[0::2, 1::2] = -M.imag
    return B


@pytest.mark.parametrize("path", CG, ids=[os.path.basename(p)[:-4] for p in CG])
def test_complex_fixtures_and_the_blockified_identity(path):
    g = load(path)
    if g["status"] != 0:
        pytest.skip("not positive definite")
    M = full_matrix(g); P = g["Perm"]
    Mp = M[np.ix_(P, P)]
    L = dense_L(g)
    assert np.abs(L @ L.conj().T - Mp).max() < 1e-11 * np.abs(Mp).max()          # the reference's factor
    assert np.abs(np.diag(L).imag).max() == 0.0 and (np.diag(L).real > 0).all()
    # chol(blockified A) == blockified(chol A): what the CUDA path computes is the reference's factor
    Lb = np.linalg.cholesky(blockify(Mp))
    assert np.abs(Lb - blockify(L)).max() < 1e-12 * np.abs(L).max()
    x = np.linalg.solve(M, g["b"])
    assert np.abs(x - g["x"]).max() < 1e-10 * np.abs(x).max()


@pytest.mark.parametrize("zomplex", [False, True])
def test_blockify_csc_matches_numpy(zomplex):
    """The library's host-side blockification (no GPU needed) against the dense definition, symmetric-lower and unsymmetric."""
    from suitesparse_b200 import cholmod_host as H
    lib = C.CDLL(B200_LIB)
    lib.ssb200_debug_blockify.restype = C.c_int64
    lib.ssb200_debug_blockify.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(42)
    n = 9
    R = sp.random(n, n, density=0.4, random_state=rng) + sp.random(n, n, density=0.4, random_state=rng) * 1j + sp.eye(n) * (3 + 0.7j)
    for lower in (1, 0):
        A = (sp.tril(R) if lower else R).tocsc(); A.sort_indices()
        keep = []
        class Holder: _keep = keep
        S = H.Cholmod.sparse(Holder, A, -1 if lower else 0, zomplex=zomplex)
        cnt = lib.ssb200_debug_blockify(C.byref(S), lower, None, None, None, 0)
        p2 = np.zeros(2 * n + 1, np.int64); i2 = np.zeros(cnt, np.int64); x2 = np.zeros(cnt)
        assert lib.ssb200_debug_blockify(C.byref(S), lower, p2.ctypes.data, i2.ctypes.data, x2.ctypes.data, cnt) == cnt
        B = sp.csc_matrix((x2, i2, p2), shape=(2 * n, 2 * n)).toarray()
        D = A.toarray()
        if lower:
            D = D - 1j * np.diag(np.diag(D).imag)                # the imaginary part of the diagonal is dropped
            want = np.tril(blockify(D))
        else:
            want = blockify(D)
        assert np.array_equal(B, want)


def _run_gpu(g, zomplex=False):
    from suitesparse_b200 import cholmod_host as H
    ch = H.Cholmod(gpu=True)
    n = int(g["n"])
    U = sp.csc_matrix((g["Ux"], g["Ui"], g["Up"]), shape=(n, int(g["ncolU"])))
    S = ch.sparse(U, int(g["stype"]), zomplex=zomplex)
    ch.cm.supernodal = H.CHOLMOD_SUPERNODAL
    L = ch.analyze(S, g["perm"]) if g["perm"].size else ch.lib.cholmod_l_analyze(C.byref(S), C.byref(ch.cm))
    ok = ch.factorize(S, L, beta=(float(g["beta"]) if g["stype"] == 0 else None))
    return ch, L, ok


@pytest.mark.gpu
@pytest.mark.parametrize("path", CG, ids=[os.path.basename(p)[:-4] for p in CG])
def test_complex_dropin_against_reference(path):
    """cholmod_l_factorize / cholmod_l_solve on complex and zomplex input through the interposed symbols: L->x in CHOLMOD's
    complex layout, minor, status and the solution equal the reference's."""
    from conftest import persuper_relerr
    from suitesparse_b200 import cholmod_host as H
    g = load(path)
    ch, L, ok = _run_gpu(g, zomplex=bool(g["zomplex"]))
    assert ok == 1 and ch.cm.status == int(g["status"]) and ch.cm.gpuNumKernelLaunches > 0 and ch.cm.cpu_potrf_calls == 0
    f = ch.factor_arrays(L)
    assert f["xtype"] == H.CHOLMOD_COMPLEX and f["minor"] == int(g["minor"])
    for k in ("super", "pi", "px", "s"):
        assert np.array_equal(f[k], g[k])
    if g["status"] == 0:
        assert persuper_relerr(g["px"], f["x"], g["Lx"]) < 1e-11
        x = ch.solve(L, g["b"])
        assert np.abs(x - g["x"]).max() < 1e-9 * np.abs(g["x"]).max()
        M = full_matrix(g)
        assert np.linalg.norm(M @ x - g["b"]) / np.linalg.norm(g["b"]) < 1e-10
        # several right-hand sides; a factor changed by the caller (L -> 2L) is taken from the host again
        B = np.stack([g["b"], 1j * g["b"], g["b"].conj()], axis=1)
        X = ch.solve(L, B)
        assert np.abs(X[:, 1] - 1j * X[:, 0]).max() < 1e-10 * np.abs(X).max()
        f["x"][:] *= 2.0
        ch.b200.ssb200_invalidate_factor.argtypes = [C.c_void_p]
        ch.b200.ssb200_invalidate_factor(L)
        x2 = ch.solve(L, g["b"])
        assert np.abs(x2 - x / 4).max() < 1e-10 * np.abs(x).max()
    else:
        # not positive definite: everything before the failing supernode equals the reference, the rest follows its protocol
        assert persuper_relerr(g["px"], f["x"], g["Lx"]) < 1e-11
    ch.free_factor(L)
