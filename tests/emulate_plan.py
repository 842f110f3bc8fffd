"""numpy replay of the host plan (ssb_plan.cpp) — TEST INFRASTRUCTURE.  Executes one rank's launch list job by job on
the CPU (no tiles, no CUDA), optionally with broadcasts between ranks, so that the update enumeration, the etree shard,
the panel-cyclic distribution and the step/broadcast order can be checked against the oracle without a GPU."""
import ctypes as C
import numpy as np
from conftest import B200_LIB

L_GEMM_BIG, L_GEMM_SMALL, L_POTRF, L_TRSM, L_TRSM_TC = range(5)
NB_MID = 256


def export_plan(n, super_, pi, px, s, nranks=1, rank=0):
    lib = C.CDLL(B200_LIB)
    a = [np.ascontiguousarray(v, dtype=np.int64) for v in (super_, pi, px, s)]
    sizes = np.zeros(8, dtype=np.int64)
    P = lambda v: v.ctypes.data_as(C.c_void_p)
    rc = lib.ssb200_export_begin(C.c_int64(n), C.c_int64(len(super_) - 1), *[P(v) for v in a], C.c_int(nranks), C.c_int(rank), P(sizes))
    assert rc == 0, rc
    nl, ng, npo, nt, ns, nu = [int(v) for v in sizes[:6]]
    out = dict(launches=np.zeros((nl, 7), np.int64), gemm=np.zeros((ng, 9), np.int64), potrf=np.zeros((npo, 6), np.int64),
               trsm=np.zeros((nt, 6), np.int64), steps=np.zeros((ns, 7), np.int64), updates=np.zeros((nu, 6), np.int64),
               owner=np.zeros(len(super_) - 1, np.int32))
    rc = lib.ssb200_export_fetch(*[P(out[k]) for k in ("launches", "gemm", "potrf", "trsm", "steps", "updates", "owner")])
    assert rc == 0
    out.update(relmap_size=int(sizes[6]), nlevels=int(sizes[7]), nranks=nranks, rank=rank)
    return out


def export_plan_compact(n, super_, pi, px, s, nranks, rank):
    """Distributed-storage plan of one rank: job offsets are LOCAL; extra keys lpx, lxsize, pieces (step, home_off, cnt)."""
    lib = C.CDLL(B200_LIB)
    a = [np.ascontiguousarray(v, dtype=np.int64) for v in (super_, pi, px, s)]
    sizes = np.zeros(8, dtype=np.int64); cs = np.zeros(2, dtype=np.int64)
    P = lambda v: v.ctypes.data_as(C.c_void_p)
    rc = lib.ssb200_export_begin_compact(C.c_int64(n), C.c_int64(len(super_) - 1), *[P(v) for v in a], C.c_int(nranks), C.c_int(rank), P(sizes), P(cs))
    assert rc == 0, rc
    nl, ng, npo, nt, ns, nu = [int(v) for v in sizes[:6]]
    lpx = np.zeros(len(super_), np.int64); pieces = np.zeros((max(int(cs[1]), 1), 5), np.int64); nxt = np.zeros(max(ns, 1), np.int32)
    lib.ssb200_export_compact_fetch.restype = C.c_int64
    nsj = lib.ssb200_export_compact_fetch(P(lpx), P(pieces), P(nxt), None, C.c_int64(0))
    solve = np.zeros((max(int(nsj), 1), 4), np.int64)
    lib.ssb200_export_compact_fetch(P(lpx), P(pieces), P(nxt), P(solve), C.c_int64(nsj))
    tr = np.zeros((len(super_) - 1, 4), np.int64)
    ring_depth = int(lib.ssb200_export_compact_tr(P(tr)))
    lib.ssb200_export_compact_deps.restype = C.c_int64
    nd = lib.ssb200_export_compact_deps(None, C.c_int64(0))
    deps = np.zeros((max(int(nd), 1), 2), np.int64)
    lib.ssb200_export_compact_deps(P(deps), C.c_int64(nd))
    out = dict(launches=np.zeros((nl, 7), np.int64), gemm=np.zeros((ng, 9), np.int64), potrf=np.zeros((npo, 6), np.int64),
               trsm=np.zeros((nt, 6), np.int64), steps=np.zeros((ns, 7), np.int64), updates=np.zeros((nu, 6), np.int64),
               owner=np.zeros(len(super_) - 1, np.int32))
    rc = lib.ssb200_export_fetch(*[P(out[k]) for k in ("launches", "gemm", "potrf", "trsm", "steps", "updates", "owner")])
    assert rc == 0
    out.update(relmap_size=int(sizes[6]), nlevels=int(sizes[7]), nranks=nranks, rank=rank, lpx=lpx, lxsize=int(cs[0]),
               pieces=pieces[:int(cs[1])], step_next=nxt[:ns], solve=solve[:int(nsj)], deps=deps[:int(nd)],
               tr=tr, ring_depth=ring_depth, px=np.asarray(px, dtype=np.int64), pi=np.asarray(pi, dtype=np.int64), super=np.asarray(super_, dtype=np.int64))
    return out


def local_of(pl, home):
    """HostPlan::local_of restated: home offset in Lx -> offset in this rank's storage (transient root: own panels packed,
    the others in ring slots)."""
    px = pl["px"]
    t = int(np.searchsorted(px, home, side="right") - 1)
    assert pl["lpx"][t] >= 0
    if pl["tr"][t, 3] > 0:                                   # trailing rows only
        nsrow = int(pl["pi"][t + 1] - pl["pi"][t]); rmin = int(pl["tr"][t, 3])
        col, row = divmod(int(home - px[t]), nsrow)
        assert row >= rmin
        return int(pl["lpx"][t] + col * (nsrow - rmin) + row - rmin)
    if not pl["tr"][t, 0]:
        return int(pl["lpx"][t] + home - px[t])
    nsrow = int(pl["pi"][t + 1] - pl["pi"][t])
    col, row = divmod(int(home - px[t]), nsrow)
    J = col // NB_MID
    within = (col - J * NB_MID) * nsrow + row
    P = NB_MID * nsrow
    off = -1 - int(pl["owner"][t]); nr = pl["nranks"]; r = pl["rank"]
    if (J + off) % nr == r:
        Jf = (r - off) % nr
        return int(pl["tr"][t, 1] + ((J - Jf) // nr) * P + within)
    return int(pl["tr"][t, 2] + (J % pl["ring_depth"]) * P + within)


def assemble_compact(plan, super_, pi, px, s, S_lower, Lx):
    """scatter_A_kernel under distributed storage: a rank assembles the columns it computes, at their LOCAL offsets."""
    nsuper = len(super_) - 1
    Sp, Si, Sx = S_lower.indptr, S_lower.indices, S_lower.data
    lpx = plan["lpx"]
    for sn in range(nsuper):
        k1, k2 = int(super_[sn]), int(super_[sn + 1])
        rows = s[pi[sn]: pi[sn + 1]]; nsrow = len(rows)
        o = plan["owner"][sn]
        for k in range(k1, k2):
            if (o != plan["rank"]) if o >= 0 else (((k - k1) // NB_MID + (-1 - o)) % plan["nranks"] != plan["rank"]):
                continue
            assert lpx[sn] >= 0
            colbase = local_of(plan, int(px[sn]) + (k - k1) * nsrow)          # the column's first entry (transient root: packed panels)
            for p in range(Sp[k], Sp[k + 1]):
                i = Si[p]
                if i >= k:
                    pos = np.searchsorted(rows, i)
                    if pos < nsrow and rows[pos] == i:
                        Lx[colbase + pos] = Sx[p]


def run_lockstep_compact(plans, rel, Lx, px, selective=False):
    """Distributed storage, all ranks in one process, asynchronous pulls: a pull snapshots the source's range when the step's
    range becomes final and is delivered as LATE as the schedule allows: at the next wait_remote step, or - selective - only
    when the receiving rank reaches a step that names the pull's step as a dependency (what ssb200_mg_factorize waits for).
    A missing dependency shows up as a wrong factor.  Returns doubles pulled per rank."""
    nr = len(plans)
    if selective:
        return _run_lockstep_compact_selective(plans, rel, Lx, px)
    pending = []
    pulled = [0] * nr
    loc = local_of
    by_step = [dict() for _ in range(nr)]
    for r in range(nr):
        for k, ho, cnt, ncols, sld in plans[r]["pieces"]:
            for c in range(int(ncols)):
                by_step[r].setdefault(int(k), []).append((int(ho + c * sld), int(cnt)))
    for k in range(len(plans[0]["steps"])):
        if plans[0]["steps"][k][6]:
            for r, dst, data in pending:
                Lx[r][dst: dst + len(data)] = data
            pending = []
        srcs = set()
        for r in range(nr):
            lo, mid, hi, src, off, cnt, wait = plans[r]["steps"][k]
            run_launches(plans[r], rel, Lx[r], lo, mid)
            srcs.add((int(src), int(off), int(cnt)))
        assert len(srcs) == 1
        src = srcs.pop()[0]
        for r in range(nr):
            for ho, cnt in by_step[r].get(k, []):
                assert src >= 0 and r != src
                so = loc(plans[src], ho); dn = loc(plans[r], ho)
                pending.append((r, dn, Lx[src][so: so + cnt].copy())); pulled[r] += cnt
        for r in range(nr):
            lo, mid, hi = plans[r]["steps"][k][:3]
            run_launches(plans[r], rel, Lx[r], mid, hi)
    for r, dst, data in pending:
        Lx[r][dst: dst + len(data)] = data
    return pulled


def _run_lockstep_compact_selective(plans, rel, Lx, px):
    nr = len(plans)
    pulled = [0] * nr
    loc = local_of
    by_step = [dict() for _ in range(nr)]; deps = [dict() for _ in range(nr)]
    for r in range(nr):
        for k, ho, cnt, ncols, sld in plans[r]["pieces"]:
            for c in range(int(ncols)):
                by_step[r].setdefault(int(k), []).append((int(ho + c * sld), int(cnt)))
        for k, d in plans[r]["deps"]:
            deps[r].setdefault(int(k), []).append(int(d))
    pending = [dict() for _ in range(nr)]                 # per rank: step -> [(dst, data)]
    for k in range(len(plans[0]["steps"])):
        src = int(plans[0]["steps"][k][3])
        for r in range(nr):
            for d in deps[r].get(k, []):
                assert d < k
                for dst, data in pending[r].pop(d, []):
                    Lx[r][dst: dst + len(data)] = data
            lo, mid, hi = plans[r]["steps"][k][:3]
            run_launches(plans[r], rel, Lx[r], lo, mid)
        for r in range(nr):
            for ho, cnt in by_step[r].get(k, []):
                so = loc(plans[src], ho); dn = loc(plans[r], ho)
                pending[r].setdefault(k, []).append((dn, Lx[src][so: so + cnt].copy())); pulled[r] += cnt
        for r in range(nr):
            lo, mid, hi = plans[r]["steps"][k][:3]
            run_launches(plans[r], rel, Lx[r], mid, hi)
    for r in range(nr):
        for lst in pending[r].values():
            for dst, data in lst:
                Lx[r][dst: dst + len(data)] = data
    return pulled


def gather_compact(plans, Lx, px, xsize):
    """The host factor: every step's range copied out by its source rank (what ssb200_mg_factorize does with Lx_host)."""
    out = np.full(xsize, np.nan)
    for k in range(len(plans[0]["steps"])):
        lo, mid, hi, src, off, cnt, wait = plans[0]["steps"][k]
        if src < 0 or cnt <= 0:
            continue
        so = local_of(plans[src], int(off))
        assert np.isnan(out[off: off + cnt]).all()              # every entry exactly once
        out[off: off + cnt] = Lx[src][so: so + cnt]
    assert not np.isnan(out).any()
    return out


def relmap_of(plan, pi, s):
    rel = np.zeros(max(plan["relmap_size"], 1), dtype=np.int64)
    for d, sn, p0, nd1, nd2, moff in plan["updates"]:
        rows = s[pi[d] + p0: pi[d] + p0 + nd2]
        tgt = s[pi[sn]: pi[sn + 1]]
        pos = np.searchsorted(tgt, rows)
        assert np.all(tgt[pos] == rows)
        rel[moff: moff + nd2] = pos
    return rel


def assemble(plan, super_, pi, px, s, S_lower, Lx, beta=0.0):
    """scatter_A_kernel with the ownership rule of the shard."""
    nsuper = len(super_) - 1
    Sp, Si, Sx = S_lower.indptr, S_lower.indices, S_lower.data
    for sn in range(nsuper):
        k1, k2 = int(super_[sn]), int(super_[sn + 1])
        rows = s[pi[sn]: pi[sn + 1]]; nsrow = len(rows)
        for k in range(k1, k2):
            o = plan["owner"][sn]
            if plan["nranks"] > 1 and (o != plan["rank"] if o >= 0 else ((k - k1) // NB_MID + (-1 - o)) % plan["nranks"] != plan["rank"]):
                continue
            for p in range(Sp[k], Sp[k + 1]):
                i = Si[p]
                if i >= k:
                    pos = np.searchsorted(rows, i)
                    if pos < nsrow and rows[pos] == i:
                        Lx[px[sn] + pos + (k - k1) * nsrow] = Sx[p]
            if beta:
                Lx[px[sn] + (k - k1) * (nsrow + 1)] += beta


def run_launches(plan, rel, Lx, lo, hi):
    for kind, job0, njobs in plan["launches"][lo:hi, :3]:
        if kind in (L_GEMM_BIG, L_GEMM_SMALL):
            for a_off, c_off, moff, lda, ldc, K, nd1, nd2, c_col0 in plan["gemm"][job0: job0 + njobs]:
                idx = a_off + np.arange(nd2)[:, None] + np.arange(K)[None, :] * lda
                Pm = Lx[idx]
                Cm = Pm @ Pm[:nd1].T
                ii, jj = np.nonzero(np.arange(nd2)[:, None] >= np.arange(nd1)[None, :])
                if moff >= 0:
                    tgt = c_off + rel[moff + ii] + (rel[moff + jj] - c_col0) * ldc
                else:
                    tgt = c_off + ii + jj * ldc
                np.subtract.at(Lx, tgt, Cm[ii, jj])
        elif kind == L_POTRF:
            for x_off, lda, w, rows_below, col0, snode in plan["potrf"][job0: job0 + njobs]:
                idx = x_off + np.arange(w)[:, None] + np.arange(w)[None, :] * lda
                Bm = np.tril(Lx[idx]); Bm = Bm + np.tril(Bm, -1).T
                Lc = np.linalg.cholesky(Bm)
                ii, jj = np.tril_indices(w)
                Lx[idx[ii, jj]] = Lc[ii, jj]
        else:
            for x_off, lda, w, rows_below, col0, snode in plan["trsm"][job0: job0 + njobs]:
                idx = x_off + np.arange(w)[:, None] + np.arange(w)[None, :] * lda
                L11 = np.tril(Lx[idx])
                bidx = x_off + w + np.arange(rows_below)[:, None] + np.arange(w)[None, :] * lda
                Lx[bidx] = np.linalg.solve(L11, Lx[bidx].T).T


def run_two_streams(plan, rel, Lx, prefer):
    """Look-ahead schedule on two in-order streams with events (stream 0 = main, 1 = panel chain): a launch runs when its
    stream has reached it and its wait event has been recorded.  `prefer` picks which stream advances whenever both can -
    the two extremes (main stream as early as possible / as late as possible) expose a missing dependency as a wrong
    factor.  Returns the number of launches that ran while the OTHER stream still had earlier list entries pending
    (i.e. how much reordering the schedule allowed)."""
    L = plan["launches"]
    queues = [[t for t in range(len(L)) if L[t, 4] == st] for st in (0, 1)]
    pos = [0, 0]; recorded = set(); reordered = 0
    while pos[0] < len(queues[0]) or pos[1] < len(queues[1]):
        ready = []
        for st in (0, 1):
            if pos[st] < len(queues[st]):
                t = queues[st][pos[st]]
                if L[t, 5] < 0 or int(L[t, 5]) in recorded:
                    ready.append(st)
        assert ready, "deadlock in the look-ahead schedule"
        st = prefer if prefer in ready else ready[0]
        t = queues[st][pos[st]]
        other = 1 - st
        if pos[other] < len(queues[other]) and queues[other][pos[other]] < t:
            reordered += 1
        run_launches(plan, rel, Lx, t, t + 1)
        if L[t, 6] >= 0:
            recorded.add(int(L[t, 6]))
        pos[st] += 1
    return reordered


def factorize_emulated(n, super_, pi, px, s, S_lower, nranks=1, rank=0, bcast=None, beta=0.0):
    """bcast(buf_view, src) replicates a finished Lx range (None for a single rank)."""
    plan = export_plan(n, super_, pi, px, s, nranks, rank)
    rel = relmap_of(plan, pi, s)
    Lx = np.zeros(int(px[-1]))
    assemble(plan, super_, pi, px, s, S_lower, Lx, beta)
    for lo, mid, hi, src, off, cnt, wait in plan["steps"]:
        run_launches(plan, rel, Lx, lo, mid)
        if src >= 0 and bcast is not None:
            bcast(Lx[off: off + cnt], int(src))          # synchronous here; the lock-step test models the asynchrony
        run_launches(plan, rel, Lx, mid, hi)
    return Lx, plan


def run_lockstep(plans, rel, Lx):
    """All ranks in one process, with ASYNCHRONOUS broadcast semantics: a broadcast is snapshotted when it starts and
    only delivered when a later step declares wait_remote (or at the end) — reading remote data too early shows up as a
    wrong factor.  Returns the number of doubles broadcast."""
    nr = len(plans)
    pending, total = [], 0
    for k in range(len(plans[0]["steps"])):
        infos = set()
        if plans[0]["steps"][k][6]:
            for src, off, data in pending:
                for r in range(nr):
                    if r != src:
                        Lx[r][off: off + len(data)] = data
            pending = []
        for r in range(nr):
            lo, mid, hi, src, off, cnt, wait = plans[r]["steps"][k]
            run_launches(plans[r], rel, Lx[r], lo, mid)
            infos.add((int(src), int(off), int(cnt), int(wait)))
        assert len(infos) == 1, infos                      # every rank issues the same collective
        src, off, cnt, wait = infos.pop()
        if src >= 0:
            pending.append((src, off, Lx[src][off: off + cnt].copy())); total += cnt
        for r in range(nr):
            lo, mid, hi = plans[r]["steps"][k][:3]
            run_launches(plans[r], rel, Lx[r], mid, hi)
    for src, off, data in pending:
        for r in range(nr):
            if r != src:
                Lx[r][off: off + len(data)] = data
    return total
