"""torchrun --nproc-per-node N scripts/dist_check.py <kind> <N> [reps]: sharded factorization vs the oracle (small) and timing."""
import os, sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from suitesparse_b200 import gen
from suitesparse_b200.cholmod_host import Cholmod, _np_view
from suitesparse_b200.dist import ShardedFactor

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)     # broadcasts must preempt queued GEMM tiles
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
kind, N = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ch = Cholmod(gpu=True)
A, p = gen.make_problem(kind, N)
S = ch.sparse(A, +1); L = ch.analyze(S, p); f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ch.factor_arrays(L).items()}
S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
Ap = _np_view(s2.p, n + 1, np.int64).copy(); Ai = _np_view(s2.i, int(Ap[n]), np.int64).copy(); Ax = _np_view(s2.x, int(Ap[n]), np.float64).copy()
Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
t0 = time.time()
sf = ShardedFactor(n, f["super"], f["pi"], f["px"], f["s"], local)
tplan = time.time() - t0
sf.upload_A(Sl)
nb = sum(1 for s_ in sf.steps if s_[0] >= 0); bbytes = sum(s_[2] for s_ in sf.steps if s_[0] >= 0) * 8
print(f"[rank {rank}] plan {tplan:.2f}s steps {len(sf.steps)} bcasts {nb} ({bbytes/1e9:.2f} GB) my_flops {sf.my_flops:.3e} of {sf.total_flops:.3e} ({100*sf.my_flops/sf.total_flops:.1f}%)", flush=True)
times = []
for it in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sf.stream):
        e0.record()
    st, minor = sf.factorize_resident()
    with torch.cuda.stream(sf.stream):
        e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(t.item())
    if it == reps - 1 and sf.phase_event is not None:
        print(f"[rank {rank}] total {e0.elapsed_time(e1):.1f} ms, subtree phase {e0.elapsed_time(sf.phase_event):.1f} ms, top phase {sf.phase_event.elapsed_time(e1):.1f} ms", flush=True)
if os.environ.get("SSB200_DIST_STEPTIME"):
    ev_full = sf.step_events
    dt_full = np.array([ev_full[k].elapsed_time(ev_full[k + 1]) for k in range(len(ev_full) - 1)])
    os.environ["SSB200_DIST_DEBUG"] = "nocomm"
    sf.factorize_resident(); torch.cuda.synchronize()
    ev_nc = sf.step_events
    dt_nc = np.array([ev_nc[k].elapsed_time(ev_nc[k + 1]) for k in range(len(ev_nc) - 1)])
    os.environ["SSB200_DIST_DEBUG"] = ""
    if rank == 0:
        st = np.array(sf.steps)
        cyc = (st[:, 3] == 1)
        print("steps", len(st), "sum full %.1f nocomm %.1f" % (dt_full.sum(), dt_nc.sum()))
        # group: consecutive runs of wait_remote steps by size of bcast
        big = np.argsort(dt_full - dt_nc)[::-1][:25]
        for k in sorted(big):
            print("  step %4d src %2d cnt %10d wait %d  full %.3f ms  nocomm %.3f ms" % (k, st[k, 0], st[k, 2], st[k, 3], dt_full[k], dt_nc[k]))
        lo = int(np.argmax(cyc)) if cyc.any() else len(st)
        print("subtree part: full %.1f nocomm %.1f | top part: full %.1f nocomm %.1f" % (dt_full[:lo].sum(), dt_nc[:lo].sum(), dt_full[lo:].sum(), dt_nc[lo:].sum()))
        # top part by bcast size class
        for name, sel in (("panel steps cnt<2e6", (st[:, 2] < 2e6) & cyc), ("2e6..5e6", (st[:, 2] >= 2e6) & (st[:, 2] < 5e6) & cyc), (">=5e6", (st[:, 2] >= 5e6) & cyc)):
            print("  %-22s n=%4d full %.1f nocomm %.1f" % (name, sel.sum(), dt_full[sel].sum(), dt_nc[sel].sum()))
fl = ch.cm.fl
if rank == 0:
    print(f"{kind}{N} world={world} status={st} minor={minor} ms={['%.1f' % v for v in times]} best GF/s={fl / min(times) / 1e6:.1f}", flush=True)
# correctness: residual of a replicated solve on every rank (+ oracle on small problems)
b = np.ones(n)
y = sf.solve(b[f["Perm"]], which=2)
x = np.empty(n); x[f["Perm"]] = y
Af = A + sp.triu(A, 1).T
res = np.linalg.norm(Af @ x - b) / np.linalg.norm(b)
msg = f"[rank {rank}] resid {res:.2e}"
if n <= 40000:
    from oracle import oracle
    sto, mo, Lo = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], Sl)
    Lx = sf.download_L()
    worst = 0.0
    for s_ in range(f["nsuper"]):
        a, b_ = int(f["px"][s_]), int(f["px"][s_ + 1])
        worst = max(worst, np.abs(Lx[a:b_] - Lo[a:b_]).max() / max(np.abs(Lo[a:b_]).max(), 1e-300))
    msg += f" max persuper relerr vs oracle {worst:.2e}"
print(msg, flush=True)
dist.barrier(); dist.destroy_process_group()
