"""First GPU bring-up: plain layer vs the C oracle, then the drop-in path vs the reference CPU library."""
import sys, time, numpy as np, scipy.sparse as sp, ctypes as C
sys.path.insert(0, ".")
from suitesparse_b200 import gen
from suitesparse_b200.cholmod_host import Cholmod, _np_view
from suitesparse_b200 import plain
from oracle import oracle

ch = Cholmod(gpu=True)

def lower_of(ch, S, L):
    S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
    Ap = _np_view(s2.p, n + 1, np.int64).copy(); Ai = _np_view(s2.i, int(Ap[n]), np.int64).copy(); Ax = _np_view(s2.x, int(Ap[n]), np.float64).copy()
    ch.free_sparse(S2)
    return sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))

def persuper_err(f, Lx, Lref):
    worst = (0, -1)
    for s in range(f["nsuper"]):
        a, b = int(f["px"][s]), int(f["px"][s + 1])
        e = np.abs(Lx[a:b] - Lref[a:b]).max() / max(np.abs(Lref[a:b]).max(), 1e-300)
        if not (e <= worst[0]): worst = (e, s)
    return worst

cases = [("lap7", 4), ("lap7", 8), ("lap7", 16), ("lap27", 12), ("elas", 6), ("lap7", 32), ("lap27", 24), ("lap7", 48)]
if len(sys.argv) > 1: cases = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for kind, N in cases:
    A, p = gen.make_problem(kind, N)
    S = ch.sparse(A, +1)
    L = ch.analyze(S, p)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ch.factor_arrays(L).items()}
    Al = lower_of(ch, S, L)
    n = f["n"]
    t = time.time(); st_o, minor_o, Lo = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], Al); to = time.time() - t
    pl = plain.Plan(n, f["super"], f["pi"], f["px"], f["s"])
    up = oracle.enumerate_updates(n, f["super"], f["pi"], f["s"])
    st, minor, Lx = pl.factorize(Al)
    err, sw = persuper_err(f, Lx, Lo)
    stt = pl.stats()
    print(f"{kind}{N}: n={n} nsuper={f['nsuper']} levels={stt['nlevels']} updates={stt['nupdates']}/{len(up['d'])} status={st}/{st_o} minor={minor}/{minor_o} "
          f"max persuper relerr={err:.2e} (s={sw}) nan={np.isnan(Lx).sum()} ms_total={stt['ms_total']:.2f} (asm {stt['ms_assemble']:.2f} upd {stt['ms_update']:.2f} fac {stt['ms_factor']:.2f}) launches={stt['kernel_launches']} oracle={to:.2f}s", flush=True)
    if err > 1e-10:
        lev = np.zeros(f["nsuper"], dtype=np.int32)
        bad = [s for s in range(f["nsuper"]) if np.abs(Lx[f["px"][s]:f["px"][s+1]] - Lo[f["px"][s]:f["px"][s+1]]).max() > 1e-9 * np.abs(Lo).max()]
        print("   bad supernodes:", bad[:20], "count", len(bad))
        for s in bad[:3]:
            a, b = int(f["px"][s]), int(f["px"][s+1]); nsrow = int(f["pi"][s+1]-f["pi"][s]); nscol = int(f["super"][s+1]-f["super"][s])
            D = (Lx[a:b]-Lo[a:b]).reshape((nsrow, nscol), order="F")
            ij = np.argwhere(np.abs(D) > 1e-9*np.abs(Lo).max())[:8]
            print("    s", s, "nsrow", nsrow, "nscol", nscol, "first bad (row,col):", ij.tolist())
    # solves on the device vs oracle
    b = np.ones(n) + np.arange(n) / n
    y = pl.solve(b[f["Perm"]], which=2)
    yo = oracle.lsolve(f["super"], f["pi"], f["px"], f["s"], Lo, b[f["Perm"]]); yo = oracle.lsolve(f["super"], f["pi"], f["px"], f["s"], Lo, yo, transpose=True)
    x = np.empty(n); x[f["Perm"]] = y
    Af = A + sp.triu(A, 1).T
    print(f"   solve: |y-yo|/|yo|={np.abs(y-yo).max()/np.abs(yo).max():.2e} resid={np.linalg.norm(Af@x-b)/np.linalg.norm(b):.2e} solve_ms={pl.stats()['ms_total']:.3f}", flush=True)
    Y3 = pl.solve(np.stack([b, 2*b, b*b], axis=1)[f["Perm"], :], which=2)
    print(f"   nrhs=3 consistency: {np.abs(Y3[:,0]-y).max():.2e} {np.abs(Y3[:,1]-2*y).max():.2e}")
    pl.close()
    # drop-in path: cholmod_l_factorize -> interposed super_numeric
    ok = ch.factorize(S, L)
    f2 = ch.factor_arrays(L)
    e2, s2 = persuper_err(f, f2["x"], Lo)
    xs = ch.solve(L, b)
    print(f"   drop-in: ok={ok} status={ch.cm.status} minor={f2['minor']} relerr={e2:.2e} gpu_calls={ch.cm.gpu_syrk_calls} launches={ch.cm.gpuNumKernelLaunches} resid={np.linalg.norm(Af@xs-b)/np.linalg.norm(b):.2e}", flush=True)
    ch.free_factor(L)
print("DONE")
