"""Multi-GPU inside the library (ssb200_mg_*): timing of one mesh problem on the first ndev devices.
  python scripts/mg_bench.py [kind] [N] [ndev] [steps] [host: 0|1]"""
import sys, os, time, json
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suitesparse_b200 import gen, plain
from suitesparse_b200.cholmod_host import Cholmod, _np_view

kind = sys.argv[1] if len(sys.argv) > 1 else "lap7"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else plain._lib().ssb200_device_count()
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
with_host = len(sys.argv) > 5 and sys.argv[5] == "1"
t0 = time.time()
A, perm = gen.make_problem(kind, N)
ch = Cholmod(gpu=True)
S = ch.sparse(A, +1); Lp = ch.analyze(S, perm); fl = ch.cm.fl
f = ch.factor_arrays(Lp); n = int(f["n"])
S2 = ch.lower_permuted(S, Lp); s2 = S2.contents
Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
t_setup = time.time() - t0
t0 = time.time()
mg = plain.MultiGpu(n, f["super"], f["pi"], f["px"], f["s"], ndev=ndev)
t_plan = time.time() - t0
host = None
if with_host:
    import torch
    host_t = torch.empty(mg.xsize, dtype=torch.float64, pin_memory=False)
    host = host_t.numpy()
    t0 = time.time(); mg.pin_host(host); t_pin = time.time() - t0
st, minor = mg.factorize(Sl, Lx_host=host)
assert st == 0, (st, minor)
ms = []
for _ in range(steps):
    t0 = time.perf_counter(); st, minor = mg.factorize(Sl, Lx_host=host); ms.append((time.perf_counter() - t0) * 1e3)
info = mg.info()
if os.environ.get("SSB200_MG_TRACE"):
    tr, stp = mg.trace()
    np.savez(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"mg_trace_{kind}{N}_n{ndev}.npz"), t=tr, steps=stp)
    # compact summary: where the ranks are when the top phase starts, and the end
    first_top = int(np.argmax(stp[:, 1] > 0))
    print("trace: first top step", first_top, "reached at ms", [round(float(v), 1) for v in tr[:, first_top]], "end", [round(float(v), 1) for v in tr[:, -1]], flush=True)
b = np.ones(n)
y = mg.solve(b[f["Perm"]], which=2)
t0 = time.perf_counter(); y = mg.solve(b[f["Perm"]], which=2); t_solve = (time.perf_counter() - t0) * 1e3
info["ms_solve"] = mg.info()["ms_solve"]
x = np.empty(n); x[f["Perm"]] = y
Af = A + sp.triu(A, 1).T
resid = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
out = {"kind": kind, "N": N, "ndev": ndev, "host_copy": with_host, "ms_wall": [round(v, 1) for v in ms], "ms_internal": round(info["ms_factorize"], 1), "ms_device": round(info["ms_device"], 1),
       "tflops": round(fl / min(ms) / 1e9, 2), "resid": resid, "solve_ms_wall": round(t_solve, 1), "solve_ms_internal": round(info["ms_solve"], 1),
       "nvlink_GB": round(info["nvlink_bytes"] / 1e9, 2), "device_GB": [round(v / 1e9, 1) for v in info["device_bytes"]],
       "flop_share": [round(v / sum(info["rank_flops"]), 3) for v in info["rank_flops"]], "xsize_GB": round(mg.xsize * 8 / 1e9, 1),
       "setup_s": round(t_setup, 1), "plan_s": round(t_plan, 1)}
if with_host:
    Lx1 = host[::100003].copy()
    out["host_finite"] = bool(np.isfinite(Lx1).all())
print(json.dumps(out), flush=True)
mg.close()
