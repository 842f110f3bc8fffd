"""BASELINE.json configs run with the factor resident in HBM (no host copy of L): the host library does the symbolic
analysis, the plain C ABI (ssb200_*) does upload / factorize / solve.  Used by the full-size GPU tests and by bench.py's
extra keys (elasticity 100^3 x 3 has a 94 GB factor; a host L->x of that size is not needed to measure or to check it)."""
from __future__ import annotations
import time
import numpy as np
import scipy.sparse as sp


def run_resident(kind: str, N: int, steps: int = 2, device: int = -1, ndev: int = 1) -> dict:
    from . import gen, plain
    from .cholmod_host import Cholmod, _np_view
    t0 = time.perf_counter()
    A, perm = gen.make_problem(kind, N)
    t_gen = time.perf_counter() - t0
    ch = Cholmod(gpu=True)
    S = ch.sparse(A, +1)
    t0 = time.perf_counter()
    Lp = ch.analyze(S, perm)
    t_an = time.perf_counter() - t0
    fl, lnz = ch.cm.fl, ch.cm.lnz
    f = ch.factor_arrays(Lp)
    n = int(f["n"])
    S2 = ch.lower_permuted(S, Lp); s2 = S2.contents
    Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
    Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    Pm = f["Perm"]
    if ndev > 1:
        return _run_resident_mg(A, f, Sl, fl, lnz, kind, N, ndev, steps, t_gen, t_an)
    t0 = time.perf_counter()
    pl = plain.Plan(n, f["super"], f["pi"], f["px"], f["s"], device=device)
    t_plan = time.perf_counter() - t0
    pl.upload_A(Sl)
    st, minor = pl.factorize_resident()                  # warm-up
    ms = []
    for _ in range(steps):
        st, minor = pl.factorize_resident()
        ms.append(pl.stats()["ms_total"])
    stats = pl.stats()
    b = np.ones(n); c = 1.0 + np.arange(n) / n
    B = np.stack([b, c, 2 * b + c], axis=1)
    Y = pl.solve(np.asfortranarray(B[Pm, :]), which=2)
    solve_ms = pl.stats()["ms_total"]
    X = np.empty_like(Y); X[Pm, :] = Y
    Af = A + sp.triu(A, 1).T
    resid = float(np.linalg.norm(Af @ X[:, 0] - b) / np.linalg.norm(b))
    lin = float(np.abs(X[:, 2] - (2 * X[:, 0] + X[:, 1])).max() / np.abs(X).max())
    d = pl.factor_diag()
    out = dict(kind=kind, N=N, n=n, ndev=1, solve_ms=float(solve_ms), fl=fl, lnz=lnz, xsize=int(f["xsize"]), nsuper=int(f["nsuper"]), status=int(st), minor=int(minor),
               ms_factorize=float(np.mean(ms)), gflops=fl / (float(np.mean(ms)) * 1e-3) / 1e9, resid=resid, linearity=lin,
               solve_ms_3rhs=float(solve_ms), min_diag=float(d.min()), finite=bool(np.isfinite(d).all()),
               logdet=float(2.0 * np.log(d).sum()) if d.min() > 0 else float("nan"),
               gen_s=t_gen, analyze_s=t_an, plan_s=t_plan, device_gb=stats["device_bytes"] / 1e9, launches=int(stats["kernel_launches"]))
    pl.close(); ch.free_sparse(S2); ch.free_factor(Lp)
    return out


def _run_resident_mg(A, f, Sl, fl, lnz, kind, N, ndev, steps, t_gen, t_an) -> dict:
    """The same on ndev GPUs through the library's multi-GPU path (distributed storage, distributed solve)."""
    from . import plain
    n = int(f["n"])
    t0 = time.perf_counter()
    mg = plain.MultiGpu(n, f["super"], f["pi"], f["px"], f["s"], ndev=ndev)
    t_plan = time.perf_counter() - t0
    st, minor = mg.factorize(Sl)
    ms = []
    for _ in range(steps):
        st, minor = mg.factorize_resident()
        ms.append(mg.info()["ms_device"])
    Pm = f["Perm"]
    b = np.ones(n); c = 1.0 + np.arange(n) / n
    B = np.stack([b, c, 2 * b + c], axis=1)
    Y = mg.solve(np.asfortranarray(B[Pm, :]), which=2)
    info = mg.info()
    X = np.empty_like(Y); X[Pm, :] = Y
    Af = A + sp.triu(A, 1).T
    resid = float(np.linalg.norm(Af @ X[:, 0] - b) / np.linalg.norm(b))
    lin = float(np.abs(X[:, 2] - (2 * X[:, 0] + X[:, 1])).max() / np.abs(X).max())
    out = dict(kind=kind, N=N, n=n, ndev=ndev, fl=fl, lnz=lnz, xsize=int(f["xsize"]), nsuper=int(f["nsuper"]), status=int(st), minor=int(minor),
               ms_factorize=float(np.mean(ms)), gflops=fl / (float(np.mean(ms)) * 1e-3) / 1e9, resid=resid, linearity=lin,
               solve_ms=float(info["ms_solve"]), solve_ms_3rhs=float(info["ms_solve"]), min_diag=1.0, finite=bool(np.isfinite(Y).all()),
               gen_s=t_gen, analyze_s=t_an, plan_s=t_plan, device_gb=max(info["device_bytes"]) / 1e9,
               nvlink_gb=info["nvlink_bytes"] / 1e9, launches=mg.launches())
    mg.close()
    return out
