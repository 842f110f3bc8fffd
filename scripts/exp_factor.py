"""Tuning aid (GPU box): resident factorization timing of one mesh problem under the current environment switches.
  python scripts/exp_factor.py [kind] [N] [steps]      prints ms/step, TF/s, per-kernel times when the look-ahead is off."""
import sys, os, time, json
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suitesparse_b200 import gen, plain
from suitesparse_b200.cholmod_host import Cholmod, _np_view

kind = sys.argv[1] if len(sys.argv) > 1 else "lap7"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tag = os.environ.get("EXP_TAG", "")
A, perm = gen.make_problem(kind, N)
ch = Cholmod(gpu=True)
S = ch.sparse(A, +1); Lp = ch.analyze(S, perm); fl = ch.cm.fl
f = ch.factor_arrays(Lp); n = int(f["n"])
S2 = ch.lower_permuted(S, Lp); s2 = S2.contents
Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
pl = plain.Plan(n, f["super"], f["pi"], f["px"], f["s"])
pl.upload_A(Sl)
pl.factorize_resident()
ms = []
for _ in range(steps):
    pl.factorize_resident(); ms.append(pl.stats()["ms_total"])
st = pl.stats()
out = {"tag": tag, "kind": kind, "N": N, "ms": round(float(np.mean(ms)), 2), "tflops": round(fl / np.mean(ms) / 1e9, 2),
       "ms_kind": [round(v, 2) for v in st["ms_kind"]], "launches_kind": st["launches_kind"]}
b = np.ones(n)
y = pl.solve(b[f["Perm"]], which=2); x = np.empty(n); x[f["Perm"]] = y
Af = A + sp.triu(A, 1).T
out["resid"] = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b)); out["solve_ms"] = round(pl.stats()["ms_total"], 3)
print(json.dumps(out), flush=True)
