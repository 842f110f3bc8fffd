import glob, os, sys
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "*.npz")))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libcholmod_ref.so")
B200_LIB = os.path.join(REPO, "suitesparse_b200", "csrc", "libsuitesparse_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    return {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}


def golden_matrix(g):
    import scipy.sparse as sp
    n = int(g["n"])
    S = sp.csc_matrix((g["Sx"], g["Si"], g["Sp"]), shape=(n if g["stype"] != 0 else n, int(g["ncolS"])))
    F = None
    if "Fp" in g:
        F = sp.csc_matrix((g["Fx"], g["Fi"], g["Fp"]), shape=(int(g["ncolS"]), n))
    return S, F


@pytest.fixture(scope="session")
def have_ref():
    return os.path.exists(REF_LIB)


def persuper_relerr(px, Lx, Lref):
    worst = 0.0
    for s in range(len(px) - 1):
        a, b = int(px[s]), int(px[s + 1])
        den = max(np.abs(Lref[a:b]).max(), 1e-300)
        worst = max(worst, float(np.abs(Lx[a:b] - Lref[a:b]).max() / den))
    return worst
