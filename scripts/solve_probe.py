"""Tuning aid: one factorization, then forward+backward solves (for ncu launch lists / captures of the solve kernels)."""
import sys, os, numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suitesparse_b200 import gen, plain
from suitesparse_b200.cholmod_host import Cholmod, _np_view
kind = sys.argv[1] if len(sys.argv) > 1 else "lap7"; N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
A, perm = gen.make_problem(kind, N)
ch = Cholmod(gpu=True)
S = ch.sparse(A, +1); Lp = ch.analyze(S, perm)
f = ch.factor_arrays(Lp); n = int(f["n"])
S2 = ch.lower_permuted(S, Lp); s2 = S2.contents
Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
pl = plain.Plan(n, f["super"], f["pi"], f["px"], f["s"])
pl.upload_A(sp.csc_matrix((Ax, Ai, Ap), shape=(n, n)))
pl.factorize_resident()
b = np.ones(n)
for _ in range(reps):
    y = pl.solve(b[f["Perm"]], which=2)
    print("solve ms", round(pl.stats()["ms_total"], 3), "launches", pl.stats()["kernel_launches"], flush=True)
x = np.empty(n); x[f["Perm"]] = y
Af = A + sp.triu(A, 1).T
print("resid", float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b)))
