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
print("fl %.4g ms_total %.1f  asm %.1f upd %.1f fac %.1f | kinds ms %s | TF/s per kind %s | launches %s" % (ch.cm.fl, s["ms_total"], s["ms_assemble"], s["ms_update"], s["ms_factor"],
      [round(v, 1) for v in s["ms_kind"]], [round(f / max(m, 1e-9) / 1e9, 2) for f, m in zip(s["flops_kind"], s["ms_kind"])], s["launches_kind"]))
print("factorize GF/s %.1f" % (ch.cm.fl / s["ms_total"] / 1e6))
y = pl.solve(np.ones(n), which=2)
y = pl.solve(np.ones(n), which=2)
print("solve ms", pl.stats()["ms_total"], "GB/s %.1f" % (16 * pl.xsize / pl.stats()["ms_total"] / 1e6), "launches", pl.stats()["kernel_launches"])
import scipy.sparse as sp2
x = np.empty(n); x[f["Perm"]] = y
Af = A + sp.triu(A, 1).T
print("resid", np.linalg.norm(Af @ x - np.ones(n)) / np.sqrt(n))
# per-launch efficiency table of the last factorize
import ctypes as C
lib = pl.lib
lib.ssb200_debug_launches.restype = C.c_int64; lib.ssb200_debug_launches.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
pl.factorize_resident()
nl = lib.ssb200_debug_launches(pl.h, None, 0)
buf = np.zeros(nl * 7); lib.ssb200_debug_launches(pl.h, buf.ctypes.data_as(C.c_void_p), nl * 7)
T = buf.reshape(nl, 7)
print("launch table: kind phase Kbucket | n launches | tiles | GFLOP | ms | TF/s")
import collections
agg = collections.defaultdict(lambda: [0, 0, 0.0, 0.0])
for kind, phase, ntiles, njobs, flops, ms, K in T:
    kb = 0 if kind >= 2 else (64 if K <= 64 else 256 if K <= 256 else 1024 if K <= 1024 else 4096 if K <= 4096 else 99999)
    tb = 0 if kind >= 2 else (148 if ntiles <= 148 else 592 if ntiles <= 592 else 2368 if ntiles <= 2368 else 10**9)
    a = agg[(int(kind), int(phase), kb, tb)]; a[0] += 1; a[1] += ntiles; a[2] += flops; a[3] += ms
for k in sorted(agg):
    a = agg[k]
    print("  kind %d phase %d K<=%-6d tiles<=%-10d | %5d | %9d | %10.1f | %8.2f | %6.2f" % (*k, a[0], a[1], a[2] / 1e9, a[3], a[2] / max(a[3], 1e-9) / 1e9))
