"""Workload for ncu: analyze, then N resident factorizations and one solve of <kind> <N>^3 (plain layer)."""
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
from suitesparse_b200 import gen, plain
from suitesparse_b200.cholmod_host import Cholmod, _np_view
kind, N = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ch = Cholmod(gpu=True)
A, p = gen.make_problem(kind, N)
S = ch.sparse(A, +1); L = ch.analyze(S, p); f = ch.factor_arrays(L)
S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
pl = plain.Plan(n, f["super"], f["pi"], f["px"], f["s"])
pl.upload_A(sp.csc_matrix((Ax, Ai, Ap), shape=(n, n)))
for _ in range(reps):
    st, minor = pl.factorize_resident()
s = pl.stats()
print("fl", ch.cm.fl, "ms", s["ms_total"], "kinds", s["ms_kind"], "flops", s["flops_kind"], "launches", s["launches_kind"])
y = pl.solve(np.ones(n), which=2)
print("solve ms", pl.stats()["ms_total"])
