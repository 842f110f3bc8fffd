#!/usr/bin/env python
"""bench.py — supernodal Cholesky factorize GFLOP/s (fp64) on B200, plus solve GB/s (BASELINE.json's metric).

  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --steps K --warmup W  # the reference's own CPU+BLAS path on the host cores

A "step" is one numeric factorization (cholmod_l_super_numeric) of the workload matrix with a symbolic factor that
already exists (the refactorization loop of a Newton / time-stepping code: analyze once, factorize many times).
  value   factorize GFLOP/s = Common->fl / t, A and the plan resident in HBM, timed with CUDA events on the plan's stream
  e2e     the same metric through the drop-in C-ABI call cholmod_l_super_numeric(S, NULL, beta, L, Common) with HOST
          buffers: S is uploaded and L->x (xsize doubles) is copied back inside the timed region, every step
  solve   forward+backward triangular solve GB/s = 16*xsize / t  (extra keys)
  roofline  the dominant kernel (gemm_nt_sub_kernel<128>, DMMA) against a cuBLAS DGEMM peak measured in this run
  cpu_baseline  the unmodified reference library (baseline/_ref/libcholmod.so) on the host cores, bounded sample;
                `--impl reference` times it on the full workload (fewer steps when a step takes minutes)
Workload: BASELINE.json configs[1], the 3-D 7-point Laplacian 128^3 (n = 2 097 152, L = 29 GB), geometric nested
dissection (MESHND) passed as the user permutation.  The other configs are parity-test cases (tests/).
Multi-GPU (N > 1): ONE factorization sharded over the N GPUs along the elimination tree, scaling = "strong" (the matrix is
fixed), value = fl / (max over devices of the device time).  Default (--multi mg): the library's own multi-GPU path - rank 0's
process drives all N devices through the C ABI (ssb200_mg_*: independent subtrees per device, the wide top supernodes
panel-cyclic, finished ranges pulled over NVLink by peer loads, no NCCL on the data path), the other torchrun ranks only join
CPU barriers; e2e = cholmod_l_super_numeric with SSB200_DEVICES, every device copies its share of L to the host.
--multi nccl: the round-1 path, one process per GPU with NCCL broadcasts (suitesparse_b200/dist.py).
e2e at N = 1: the caller's L->x is pageable (the host library's malloc); the factor leaves through the pinned staging ring
(DESIGN.md section 4a); e2e.page_locked_ms_per_step is the same call once L->x has been page-locked.
"""
from __future__ import annotations
import argparse, ctypes as C, json, os, subprocess, sys, threading, time
import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kind", default=os.environ.get("SSB200_BENCH_KIND", "lap7"))
    ap.add_argument("--N", type=int, default=int(os.environ.get("SSB200_BENCH_N", "128")))
    ap.add_argument("--cpu-sample-N", type=int, default=0, help="mesh size of the CPU sample (0 = choose from the step count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs reported as extra keys")
    ap.add_argument("--multi", default=os.environ.get("SSB200_BENCH_MULTI", "mg"), choices=["mg", "nccl"],
                    help="N > 1: mg = the library's own multi-GPU path (one process drives all devices through the C ABI; default), "
                         "nccl = one process per GPU with torch.distributed broadcasts (suitesparse_b200/dist.py)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------------
def build_problem(ch, kind, N):
    """A (upper CSC), ND permutation, symbolic factor, S = tril(P A P') as cholmod_factorize_p builds it."""
    from suitesparse_b200 import gen
    A, perm = gen.make_problem(kind, N)
    S = ch.sparse(A, +1)
    t0 = time.time()
    Lp = ch.analyze(S, perm)
    t_an = time.time() - t0
    S2 = ch.lower_permuted(S, Lp)
    return A, perm, S, Lp, S2, t_an


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def set_blas_threads(n):
    """torchrun exports OMP_NUM_THREADS=1, which OpenBLAS honours: set the thread count of the reference's BLAS
    explicitly (and report what the library says it uses), so the CPU arm always runs on all host cores."""
    import glob, scipy
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "GOTO_NUM_THREADS"):
        os.environ[k] = str(n)
    libs = sorted(glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")))
    ob = C.CDLL(libs[0], mode=C.RTLD_GLOBAL)
    ob.scipy_openblas_set_num_threads(int(n))
    return int(ob.scipy_openblas_get_num_threads())


def time_reference_steps(kind, N, steps, warmup, budget_s=170.0):
    """The reference's own cholmod_l_super_numeric (unmodified CHOLMOD built from /root/reference + OpenBLAS on all host
    cores) on the SAME workload as our arm.  A step = one numeric refactorization into an existing numeric L (L->x is
    allocated and touched before the timed region, as in our arm).  When steps+warmup full-size steps do not fit the time
    budget, fewer full-size steps are timed (at least one, then without warm-up) instead of shrinking the matrix."""
    from suitesparse_b200.cholmod_host import Cholmod, _np_view, CHOLMOD_REAL
    threads = set_blas_threads(host_cores())
    ch = Cholmod(gpu=False)
    A, perm, S, Lp, S2, t_an = build_problem(ch, kind, N)
    fl = ch.cm.fl
    f = ch.hot("cholmod_l_super_numeric")
    beta = (C.c_double * 2)(0.0, 0.0)
    # numeric L with its pages touched (the reference allocates L->x inside the first call: cholmod_super_numeric.c:211-223)
    ch.lib.cholmod_l_change_factor(CHOLMOD_REAL, 1, 1, 1, 1, Lp, C.byref(ch.cm))
    xsize = Lp.contents.xsize
    _np_view(Lp.contents.x, xsize, np.float64)[:] = 0.0
    t_begin = time.perf_counter()
    t0 = time.perf_counter(); ok = f(S2, None, beta, Lp, C.byref(ch.cm)); t1 = time.perf_counter() - t0
    assert ok and ch.cm.status == 0
    times = []
    if t1 * (steps + warmup) <= budget_s:
        for _ in range(max(0, warmup - 1)):
            f(S2, None, beta, Lp, C.byref(ch.cm))
        warm_eff = max(1, warmup)
        for _ in range(steps):
            t0 = time.perf_counter(); ok = f(S2, None, beta, Lp, C.byref(ch.cm)); times.append(time.perf_counter() - t0)
    else:
        warm_eff = 0
        times.append(t1)
        while len(times) < steps and (time.perf_counter() - t_begin) + 1.1 * t1 <= budget_s:
            t0 = time.perf_counter(); ok = f(S2, None, beta, Lp, C.byref(ch.cm)); times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    assert ok and ch.cm.status == 0
    # solve sample
    b = np.ones(A.shape[0])
    ts = time.perf_counter(); x = ch.solve(Lp, b); ts = time.perf_counter() - ts
    import scipy.sparse as sp
    Af = A + sp.triu(A, 1).T
    resid = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
    ch.free_sparse(S2); ch.free_factor(Lp)
    return dict(gflops=fl / dt / 1e9, sec=dt, fl=fl, xsize=xsize, solve_gbs=16.0 * xsize / ts / 1e9, resid=resid, n=A.shape[0],
                steps_effective=len(times), warmup_effective=warm_eff, threads=threads, host_lib=os.path.relpath(ch.lib._name, REPO))


class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index; self.samples = []; self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm_sorted = sorted(sm)
        # "under load": the upper half of the samples (idle gaps between steps pull the clock down)
        med = sm_sorted[len(sm_sorted) // 2] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": (max(mx) if mx else None), "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM ceiling on this GPU (MEASURED_PEAKS.json holds only bf16 and HBM)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize(dev)
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main_sharded(args, torch, dist, dev, rank, world, local, workload):
    """N > 1: one factorization sharded over the ranks (strong scaling)."""
    import scipy.sparse as sp
    from suitesparse_b200.cholmod_host import Cholmod, _np_view
    from suitesparse_b200.dist import ShardedFactor
    ch = Cholmod(gpu=True)
    A, perm, S, Lp, S2, t_an = build_problem(ch, args.kind, args.N)      # every rank analyses (deterministic, host)
    n = A.shape[0]; fl = ch.cm.fl; lnz = ch.cm.lnz
    f = ch.factor_arrays(Lp)
    xsize = int(f["xsize"]); nsuper = int(f["nsuper"])
    s2 = S2.contents
    Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
    Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    sf = ShardedFactor(n, f["super"], f["pi"], f["px"], f["s"], local)
    sf.upload_A(Sl)
    sampler = ClockSampler(local); sampler.start()

    def timed(host_out=None, upload=False, shared=False):
        dist.barrier(); torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(sf.stream):
            e0.record()
        if upload:
            sf.upload_A(Sl)
        st, minor = sf.factorize_resident(host_out=host_out, host_shared=shared)
        with torch.cuda.stream(sf.stream):
            e1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert st == 0
        return t.tolist()

    for _ in range(args.warmup):
        timed()
    tdev = [timed()[0] for _ in range(args.steps)]
    t_dev = float(np.mean(tdev))
    launches_per_step = int(sf.plan.stats()["kernel_launches"])
    # e2e: S from host memory on every rank; the host L->x is one shared-memory buffer (the application = rank 0 reads it),
    # page-locked by every rank, and every rank streams the ranges it finishes over its own PCIe link, inside the timed region
    host = sf.shared_host_factor()                # collective: a tensor on every rank, or None on every rank
    shared = host is not None
    if not shared:                                # no shared buffer: rank 0 pulls everything
        host = torch.empty(xsize, dtype=torch.float64, pin_memory=True) if rank == 0 else None
    timed(host_out=host, upload=True, shared=shared)
    twall = [timed(host_out=host, upload=True, shared=shared)[1] for _ in range(args.steps)]
    t_host = float(np.mean(twall))
    clocks = sampler.stop()
    # correctness on every rank: replicated solve
    b = np.ones(n)
    y = sf.solve(b[f["Perm"]], which=2)
    x = np.empty(n); x[f["Perm"]] = y
    Af = A + sp.triu(A, 1).T
    resid = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
    solve_ms = sf.plan.stats()["ms_total"]
    if rank == 0:
        xh = host.numpy()
        # the host copy is the factor: spot-check it against the device copy
        Ld = sf.Lx[:: 100003].cpu().numpy()
        assert np.array_equal(xh[:: 100003], Ld), "streamed host copy differs from the device factor"
        nb = sum(1 for st in sf.steps if st[0] >= 0); bbytes = sum(st[2] for st in sf.steps if st[0] >= 0) * 8
        a_bytes = int(s2.nzmax) * 16 + (n + 1) * 8
        out = {"metric": "supernodal Cholesky factorize GFLOP/s (fp64)", "value": round(fl / t_dev / 1e9, 1), "unit": "GFLOP/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_dev * 1e3, 2), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "n": n, "fl": fl, "lnz": lnz, "nsuper": nsuper, "xsize": xsize,
                          "l2": "inputs_exceed_l2 (L is %.1f GB)" % (xsize * 8 / 1e9),
                          "parallelism": f"etree shard x{world}: subtrees per rank + panel-cyclic top supernodes; {nb} NCCL broadcasts, {bbytes / 1e9:.1f} GB replicated per factorization",
                          "rank0_flop_share": round(sf.my_flops / sf.total_flops, 4)},
               "e2e": {"value": round(fl / t_host / 1e9, 1), "unit": "GFLOP/s", "h2d_bytes_per_step": a_bytes * world, "d2h_bytes_per_step": xsize * 8,
                       "ms_per_step": round(t_host * 1e3, 2), "call": "ShardedFactor.upload_A + factorize_resident(host_out=L->x in " + ("shared memory, every rank copies out its share" if shared else "rank 0's pinned memory") + "), wall clock, max over ranks"},
               "gpu_launches": launches_per_step * args.steps,
               "clocks": clocks,
               "roofline": {"kernel": "gemm_nt_sub_kernel<128>", "bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                            "note": "per-kernel roofline is reported by the N=1 run; at N>1 the whole-step rate is value/N per GPU"},
               "solve": {"ms": round(solve_ms, 3), "unit": "GB/s", "value": round(16.0 * xsize / (solve_ms * 1e-3) / 1e9, 1), "note": "replicated: every rank holds all of L", "resid_2norm_rel": resid}}
        print(json.dumps(out), flush=True)
    dist.barrier(); dist.destroy_process_group()



def run_extra(kind, N, ndev=1, steps=1):
    """Another BASELINE config with the factor resident in HBM (no host copy of L), in a child process so that its plans,
    pinned buffers and 70-100 GB of HBM are gone afterwards.  Returns a small dict or {"error": ...}."""
    code = ("import json,sys;sys.path.insert(0,%r);from suitesparse_b200 import configs;"
            "print(json.dumps(configs.run_resident(%r,%d,steps=%d,ndev=%d)))" % (REPO, kind, N, steps, ndev))
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
        r = json.loads(p.stdout.strip().splitlines()[-1])
        keep = ("kind", "N", "n", "fl", "xsize", "ndev", "ms_factorize", "gflops", "resid", "solve_ms", "device_gb")
        return {k: (round(r[k], 3) if isinstance(r[k], float) and r[k] > 1e-3 else r[k]) for k in keep if k in r}
    except Exception as ex:      # noqa
        return {"kind": kind, "N": N, "ndev": ndev, "error": str(ex)[:200]}


def main_mg(args, torch, dist, dev, rank, world, local, workload):
    """N > 1 through the library's own multi-GPU path (ssb200_mg_*, what cholmod_l_super_numeric uses with SSB200_DEVICES):
    torchrun starts one process per GPU, but this path lives in ONE process - host threads, peer access over NVLink, no
    NCCL - so rank 0 drives all N devices through the C ABI and the other ranks only take part in the barriers and in the
    max-over-ranks reduction of the times."""
    import scipy.sparse as sp
    out = None
    t_dev = t_host = 0.0
    # The waiting ranks must wait on the CPU: an NCCL barrier is a kernel that spins on THEIR GPU until rank 0 arrives, and a
    # busy context of another process time-slices the device with rank 0's kernels (measured: 936 ms instead of 312 ms per
    # step).  So: NCCL for the one all-reduce of the times at the end, a gloo group for the barriers around the timed regions.
    cpu_group = dist.new_group(backend="gloo")
    barrier = lambda: dist.barrier(group=cpu_group)
    if rank == 0:
        from suitesparse_b200.cholmod_host import Cholmod, _np_view
        from suitesparse_b200 import plain
        ch = Cholmod(gpu=True)
        A, perm, S, Lp, S2, t_an = build_problem(ch, args.kind, args.N)
        n = A.shape[0]; fl = ch.cm.fl; lnz = ch.cm.lnz
        f = ch.factor_arrays(Lp)
        xsize = int(f["xsize"]); nsuper = int(f["nsuper"])
        s2 = S2.contents
        Ap = _np_view(s2.p, n + 1, np.int64); Ai = _np_view(s2.i, int(Ap[n]), np.int64); Ax = _np_view(s2.x, int(Ap[n]), np.float64)
        Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
        t0 = time.perf_counter()
        mg = plain.MultiGpu(n, f["super"], f["pi"], f["px"], f["s"], ndev=world)
        t_plan = time.perf_counter() - t0
        st, minor = mg.factorize(Sl)                      # uploads S to every device
        assert st == 0
        sampler = ClockSampler(local); sampler.start()
    barrier(); torch.cuda.synchronize(dev)
    if rank == 0:
        for _ in range(args.warmup):
            mg.factorize_resident()
        tdev = []; launches = 0
        for _ in range(args.steps):
            st, minor = mg.factorize_resident()
            tdev.append(mg.info()["ms_device"] * 1e-3); launches += mg.launches()
        t_dev = float(np.mean(tdev))
    barrier(); torch.cuda.synchronize(dev)
    if rank == 0:
        # e2e: host S in, host L->x out (page-locked once; every device copies its share out over its own PCIe link)
        host_t = torch.empty(xsize, dtype=torch.float64)
        host = host_t.numpy()
        mg.pin_host(host)
        mg.factorize(Sl, Lx_host=host)
        tw = []
        for _ in range(args.steps):
            t0 = time.perf_counter(); st, minor = mg.factorize(Sl, Lx_host=host); tw.append(time.perf_counter() - t0)
        t_host = float(np.mean(tw))
        clocks = sampler.stop()
        info = mg.info()
        b = np.ones(n)
        mg.solve(b[f["Perm"]], which=2)
        y = mg.solve(b[f["Perm"]], which=2)
        solve_ms = mg.info()["ms_solve"]
        x = np.empty(n); x[f["Perm"]] = y
        Af = A + sp.triu(A, 1).T
        resid = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
        # the host copy is the factor: solve with it through a fresh upload of a sample? cheap check: finite and equal to a re-download
        assert np.isfinite(host[:: 100003]).all()
    barrier()
    tt = torch.tensor([t_dev, t_host], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)            # NCCL sees all N ranks; nobody is computing at this point
    torch.cuda.synchronize(dev)
    t_dev, t_host = tt.tolist()
    if rank == 0:
        a_bytes = int(s2.nzmax) * 16 + (n + 1) * 8
        out = {"metric": "supernodal Cholesky factorize GFLOP/s (fp64)", "value": round(fl / t_dev / 1e9, 1), "unit": "GFLOP/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_dev * 1e3, 2), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "n": n, "fl": fl, "lnz": lnz, "nsuper": nsuper, "xsize": xsize,
                          "l2": "inputs_exceed_l2 (L is %.1f GB)" % (xsize * 8 / 1e9),
                          "parallelism": f"etree shard x{world} inside the library (ssb200_mg_*: rank 0's process drives all {world} GPUs with one host thread per device, "
                                         f"distributed storage, NVLink pulls on peer pointers, no NCCL on the data path; the other torchrun ranks only join the barriers)",
                          "nvlink_GB_per_factorization": round(info["nvlink_bytes"] / 1e9, 2),
                          "hbm_GB_per_gpu": [round(v / 1e9, 1) for v in info["device_bytes"]],
                          "flop_share_per_gpu": [round(v / sum(info["rank_flops"]), 3) for v in info["rank_flops"]],
                          "plan_s": round(t_plan, 2)},
               "e2e": {"value": round(fl / t_host / 1e9, 1), "unit": "GFLOP/s", "h2d_bytes_per_step": a_bytes * world, "d2h_bytes_per_step": xsize * 8,
                       "ms_per_step": round(t_host * 1e3, 2),
                       "call": "ssb200_mg_factorize(host S, host L->x) = what cholmod_l_super_numeric runs with SSB200_DEVICES: S uploaded to every device, each device copies its share of L out, wall clock"},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "roofline": {"kernel": "gemm_nt_sub_kernel<128>", "bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                            "note": "per-kernel roofline is reported by the N=1 run; at N>1 the whole-step rate is value/N per GPU"},
               "solve": {"ms": round(solve_ms, 3), "unit": "GB/s", "value": round(16.0 * xsize / (solve_ms * 1e-3) / 1e9, 1),
                         "note": "distributed: every device solves with its own blocks, right-hand side on device 0 reached over NVLink", "resid_2norm_rel": resid}}
        mg.close()
        if not args.no_extra and world >= 2:
            # BASELINE configs[3], second half: elasticity 100^3 x 3 DOF on 2 B200 (factor resident, distributed)
            out["extra_configs"] = [run_extra("elas", 100, ndev=2)]
            if world >= 8:
                # a factor LARGER than one GPU's HBM: 7-point 208^3, L = 213 GB, distributed over the 8 devices (own supernodes,
                # trailing rows of the remote ones a rank reads, own panels of the root + a ring for the others)
                out["extra_configs"].append(run_extra("lap7", 208, ndev=8))
        print(json.dumps(out), flush=True)
    barrier(); dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{args.kind} {args.N}^3, geometric ND (MESHND) as UserPerm, supernodal LL', refactorization step"

    # -------------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        # same workload as our arm (args.kind, args.N) unless a bounded sample is asked for explicitly (--cpu-sample-N:
        # the in-line cpu_baseline of our arm)
        Ns = args.cpu_sample_N or args.N
        r = time_reference_steps(args.kind, Ns, args.steps, args.warmup)
        sample = (f"full workload {args.kind} {Ns}^3" if Ns == args.N else f"bounded sample {args.kind} {Ns}^3 of the {args.N}^3 workload") + \
                 f" (n={r['n']}, fl={r['fl']:.3e}), {r['steps_effective']} timed full-size step(s) after {r['warmup_effective']} warm-up"
        cfg = {"workload": workload if Ns == args.N else f"{args.kind} {Ns}^3 (sample of: {workload})", "n": r["n"], "fl": r["fl"], "xsize": int(r["xsize"]),
               "steps_effective": r["steps_effective"], "warmup_effective": r["warmup_effective"],
               "host_library": r["host_lib"], "blas": f"OpenBLAS (scipy-bundled), {r['threads']} threads set explicitly"}
        out = {"impl": "reference", "metric": "supernodal Cholesky factorize GFLOP/s (fp64)", "value": round(r["gflops"], 2), "unit": "GFLOP/s",
               "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["sec"] * 1e3, 2), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": cfg,
               "cpu_baseline": {"value": round(r["gflops"], 2), "unit": "GFLOP/s", "cores": r["threads"], "kind": "reference", "sample": sample,
                                "blas": f"OpenBLAS (scipy-bundled), {r['threads']} threads", "solve_GBps": round(r["solve_gbs"], 2), "resid": r["resid"]},
               "e2e": {"value": round(r["gflops"], 2), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out), flush=True)
        return

    # -------------------------------------------------------------------------------------------------------- our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ["SSB200_DEVICE"] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)  # broadcasts must preempt queued GEMM tiles
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    from suitesparse_b200.cholmod_host import Cholmod
    from suitesparse_b200 import plain
    if world > 1:
        if args.multi == "mg":
            return main_mg(args, torch, dist, dev, rank, world, local, workload)
        return main_sharded(args, torch, dist, dev, rank, world, local, workload)

    ch = Cholmod(gpu=True)
    A, perm, S, Lp, S2, t_an = build_problem(ch, args.kind, args.N)
    n = A.shape[0]; fl = ch.cm.fl; lnz = ch.cm.lnz
    Lc = Lp.contents
    xsize = Lc.xsize; nsuper = Lc.nsuper
    f_numeric = ch.hot("cholmod_l_super_numeric")
    beta = (C.c_double * 2)(0.0, 0.0)

    # ---- e2e: drop-in call with host buffers (first call allocates L->x, builds + caches the plan, pins L->x)
    t0 = time.perf_counter()
    ok = f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
    t_first = time.perf_counter() - t0
    if not ok or ch.cm.status != 0:
        raise SystemExit(f"factorize failed: status {ch.cm.status}")
    pl = plain.plan_of_factor(Lp)
    sampler = ClockSampler(local); sampler.start()
    first_staged = bool(pl.stats().get("d2h_staged", 0))
    # A step is one pass of a refactorization loop.  The default policy serves the first 32 factorizations into one L->x
    # through the staging ring and then page-locks it: the timed calls below measure that steady state (reached at once
    # with policy 2 = SSB200_PIN_HOST=2; the call that page-locks is second_call_s); the staged variant is timed afterwards.
    ch.b200.ssb200_set_pin_policy.restype = C.c_int; ch.b200.ssb200_set_pin_policy.argtypes = [C.c_int]
    if world == 1: ch.b200.ssb200_set_pin_policy(2)
    t0 = time.perf_counter()
    f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
    t_second = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 2)):                # the first two calls were warm-up steps as well
        f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
    if dist: dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    launches_e2e = 0
    for _ in range(args.steps):
        f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
        launches_e2e += ch.cm.gpuNumKernelLaunches
    torch.cuda.synchronize(dev)
    t_e2e = (time.perf_counter() - t0) / args.steps
    st_e2e = pl.stats()
    # the same call when L->x is never page-locked (SSB200_PIN_HOST=0; what the first 32 calls of the default policy get)
    t_locked = t_lock = None; locked_direct = None
    if world == 1:
        ch.b200.ssb200_set_pin_policy(0)
        t0 = time.perf_counter()
        f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
        t_lock = time.perf_counter() - t0
        ns = max(1, min(3, args.steps))
        torch.cuda.synchronize(dev); t0 = time.perf_counter()
        for _ in range(ns):
            f_numeric(S2, None, beta, Lp, C.byref(ch.cm))
        t_locked = (time.perf_counter() - t0) / ns
        locked_direct = pl.stats().get("d2h_staged", 0) == 1
        ch.b200.ssb200_set_pin_policy(-1)

    # ---- value: resident factorization (A already uploaded by the calls above), CUDA-event time from the plan
    for _ in range(args.warmup):
        pl.factorize_resident()
    if dist: dist.barrier()
    torch.cuda.synchronize(dev)
    ms_steps = []; launches = 0
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        st, minor = pl.factorize_resident()
        s_ = pl.stats()
        ms_steps.append(s_["ms_total"]); launches += s_["kernel_launches"]
    torch.cuda.synchronize(dev)
    wall_res = (time.perf_counter() - tw0) / args.steps
    if dist: dist.barrier()
    ms_step = float(np.mean(ms_steps))
    # ---- per-kernel times: a separate pass with the look-ahead off (one stream, CUDA events around every launch); with
    # the two-stream schedule kernels overlap and only the step total is meaningful
    kind_ms = np.zeros(6); kind_fl = np.zeros(6); kind_n = np.zeros(6); ser_ms = []
    pl.set_lookahead(False)
    pl.factorize_resident()
    n_ser = 2
    for _ in range(n_ser):
        pl.factorize_resident()
        s_ = pl.stats()
        ser_ms.append(s_["ms_total"]); kind_ms += np.array(s_["ms_kind"]); kind_fl += np.array(s_["flops_kind"]); kind_n += np.array(s_["launches_kind"])
    pl.set_lookahead(True)
    ms_serial = float(np.mean(ser_ms))

    # ---- solve: resident forward+backward, nrhs = 1
    b = np.ones(n)
    dX = torch.ones(n, dtype=torch.float64, device=dev)
    for _ in range(2):
        pl.solve_resident(dX.data_ptr(), 1, n, which=2)
    solve_ms = []
    for _ in range(max(3, args.steps)):
        dX.fill_(1.0)
        pl.solve_resident(dX.data_ptr(), 1, n, which=2)
        solve_ms.append(pl.stats()["ms_total"])
    solve_launches = pl.stats()["kernel_launches"]
    solve_ms_v = float(np.median(solve_ms))
    clocks = sampler.stop()
    # e2e solve through cholmod_l_solve (host B, host X; P and P' applied by the host library)
    x = ch.solve(Lp, b)                                  # first call builds the two solve graphs (L, L')
    ts = time.perf_counter(); x = ch.solve(Lp, b); t_solve_e2e = time.perf_counter() - ts
    import scipy.sparse as sp
    Af = A + sp.triu(A, 1).T
    resid = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))

    # ---- reduce over ranks (max time), aggregate
    t_dev = ms_step * 1e-3; t_host = t_e2e
    if dist:
        tt = torch.tensor([t_dev, t_host], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_host = tt.tolist()
    value = world * fl / t_dev / 1e9
    e2e_v = world * fl / t_host / 1e9

    if rank == 0:
        peak_tf = measure_fp64_peak(torch, dev)
        gi = 0   # gemm_nt_sub_kernel<128>
        ach = (kind_fl[gi] / (kind_ms[gi] * 1e-3) / 1e12) if kind_ms[gi] > 0 else 0.0
        names = ["gemm_nt_sub_kernel<128>", "gemm_nt_sub_kernel<64>", "potrf_block_kernel", "trsm_rows_kernel", "trsm_tc_kernel"]
        share = {names[k]: round(float(kind_ms[k] / max(kind_ms.sum(), 1e-9)), 4) for k in range(5)}
        a_bytes = int(S2.contents.nzmax) * 16 + (n + 1) * 8
        out = {"metric": "supernodal Cholesky factorize GFLOP/s (fp64)", "value": round(value, 1), "unit": "GFLOP/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_dev * 1e3, 2), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "n": n, "nnz_tril_A": int(S2.contents.nzmax), "fl": fl, "lnz": lnz, "nsuper": int(nsuper), "xsize": int(xsize),
                          "levels": st_e2e["nlevels"], "updates": st_e2e["nupdates"], "l2": "inputs_exceed_l2 (L is %.1f GB)" % (xsize * 8 / 1e9),
                          "parallelism": "1 GPU",
                          "analyze_s_host": round(t_an, 2), "first_call_s": round(t_first, 2), "first_call_staged": first_staged,
                          "second_call_s": round(t_second, 2),
                          "hot_path_library": os.path.relpath(ch.b200._name, REPO),
                          "host_cholmod_for_analyze_only": os.path.relpath(ch.lib._name, REPO),
                          "reference_blas_calls_during_our_steps": int(ch.cm.cpu_syrk_calls + ch.cm.cpu_gemm_calls + ch.cm.cpu_potrf_calls + ch.cm.cpu_trsm_calls)},
               "e2e": {"value": round(e2e_v, 1), "unit": "GFLOP/s", "h2d_bytes_per_step": a_bytes, "d2h_bytes_per_step": int(xsize) * 8,
                       "ms_per_step": round(t_host * 1e3, 2), "call": "cholmod_l_super_numeric(S,NULL,beta,L,Common) via the interposed C ABI, host buffers, steady state of a refactorization loop: "
                               "L->x page-locked (the default policy does that after 32 factorizations into one L->x; here at the second call, second_call_s). "
                               "staged_ms_per_step: the same call into pageable L->x through the pinned staging ring (the first 32 calls, or SSB200_PIN_HOST=0)" if world == 1 else
                               "cholmod_l_super_numeric(S,NULL,beta,L,Common) via the interposed C ABI, host buffers, L->x page-locked once",
                       "staged_ms_per_step": round(t_locked * 1e3, 2) if t_locked else None,
                       "staged": bool(st_e2e.get("d2h_staged", 0)), "staged_variant_ran_staged": locked_direct,
                       "cold_first_call_s": round(t_first, 2),
                       "ms_h2d": round(st_e2e["ms_h2d"], 2), "ms_d2h_exposed": round(st_e2e["ms_d2h"], 2), "ms_device_factorize": round(st_e2e["ms_total"], 2)},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "roofline": {"kernel": names[gi], "bound": "tensor", "achieved": round(ach, 2), "peak": round(peak_tf, 2), "unit": "TFLOP/s",
                            "frac": round(ach / peak_tf, 4) if peak_tf else None,
                            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed ncu --set full capture
                            # (the launch is the K = 1024 trailing update of the root's first outer panel: 20 301 tiles, 677.8 GFLOP)
                            "traffic": 7.88e9,
                            "traffic_note": "profiles/r2_ncu_full_gemm128_update_banded.txt: 5.28 GB read + 2.60 GB written in 20.6 ms (4.7 % of DRAM peak) for 5.51 GB "
                                            "of algorithmic bytes (0.21 GB operand panel + 5.30 GB read-modify-write of the target); the launch runs at 32.9 TFLOP/s "
                                            "with the DMMA pipe 89.4 % active.  Before the tiles were enumerated in bands of 8 tile columns the 211 MB panel was "
                                            "re-read once per tile column: 22.09 GB (profiles/r2_ncu_full_gemm128_update.txt), same duration",
                            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run; MEASURED_PEAKS.json has no fp64 entry; "
                                           "profiles/r2_fp64_peak.json: 35.44 burst / 35.45 sustained at 8192^3, 36.15 / 36.15 at 16384^3",
                            "kernel_time_share": share, "launches_per_step": [int(v / n_ser) for v in kind_n[:5]],
                            "kernel_timing": f"separate pass of {n_ser} steps with the look-ahead schedule off (one stream, events around every launch): {ms_serial:.1f} ms/step; the timed steps of `value` run the two-stream look-ahead schedule",
                            "whole_step_frac_of_peak": round(fl / t_dev / 1e12 / peak_tf, 4) if peak_tf else None,
                            "ncu_capture": "profiles/r2_ncu_full_gemm128_update_banded.txt (one K = 1024 launch), profiles/r2_launches_bench_lap7_128.txt (launch list: this kernel 92.9 % of the device time)"},
               "solve": {"value": round(16.0 * xsize / (solve_ms_v * 1e-3) / 1e9, 1), "unit": "GB/s", "ms": round(solve_ms_v, 3), "launches": int(solve_launches),
                         "hbm_peak_GBps": json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0,
                         "e2e_cholmod_l_solve_ms": round(t_solve_e2e * 1e3, 2), "resid_2norm_rel": resid},
               "wall_ms_per_resident_step": round(wall_res * 1e3, 2)}
        out["solve"]["frac_of_hbm"] = round(out["solve"]["value"] / out["solve"]["hbm_peak_GBps"], 4)
        if not args.no_cpu_baseline and world == 1:
            # bounded sample (about 10-30 s of CPU work on this box's cores): one size below the workload; the full-size
            # reference run is the `--impl reference` arm
            Ns = args.cpu_sample_N or min(args.N, 80)
            # separate process: the interposed symbols of this process must not be bound there
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-N", str(Ns),
                                    "--kind", args.kind, "--N", str(args.N)], capture_output=True, text=True, timeout=600)
                ref = json.loads(p.stdout.strip().splitlines()[-1])
                out["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as ex:          # noqa
                out["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        if not args.no_extra:
            # the other BASELINE configs that fit one B200, factor resident in HBM: configs[3] at full size, configs[2] at the
            # largest size that fits (27-point 256^3 needs 483 GB)
            ch.free_factor(Lp); ch.b200.cholmod_l_gpu_deallocate(C.byref(ch.cm))
            out["extra_configs"] = [run_extra("elas", 100), run_extra("lap27", 160)]
        print(json.dumps(out), flush=True)
    ch.free_sparse(S2)
    if dist:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
