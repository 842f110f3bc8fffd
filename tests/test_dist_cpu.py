"""CPU tests of the N>1 path (world_size 2 and 4, gloo): every rank builds ITS shard of the plan with the C++ plan
builder, replays its launch list in numpy (tests/emulate_plan.py) and exchanges the finished Lx ranges with
torch.distributed broadcasts exactly where the schedule says.  The result on every rank must be the oracle's factor."""
import os, sys
import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import GOLDEN, load_golden, golden_matrix, persuper_relerr, REPO


def _worker(rank, world, port, name, q):
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
    import emulate_plan as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden([p for p in GOLDEN if os.path.basename(p) == name + ".npz"][0])
    S, _ = golden_matrix(g)

    def bcast(view, src):
        t = torch.from_numpy(view)          # shares memory with the Lx slice
        dist.broadcast(t, src)

    Lx, plan = E.factorize_emulated(int(g["n"]), g["super"], g["pi"], g["px"], g["s"], S, nranks=world, rank=rank, bcast=bcast)
    err = persuper_relerr(g["px"], Lx, g["Lx"])
    nb = sum(1 for st in plan["steps"] if st[3] >= 0)
    mine = int((plan["owner"] == rank).sum())
    q.put((rank, err, nb, mine, int((plan["owner"] < 0).sum())))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("bcsstk01_tri_norelax", 2), ("bcsstk01_tri", 2), ("bcsstk01_tri_norelax", 4)])
def test_sharded_schedule_gloo(name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(r[0] for r in res) == list(range(world))
    for rank, err, nb, mine, ncyc in res:
        assert err < 1e-11, (rank, err)
        assert nb > 0                       # something was exchanged
    assert sum(r[3] for r in res) + res[0][4] == len(load_golden([p for p in GOLDEN if os.path.basename(p) == name + ".npz"][0])["super"]) - 1


@pytest.mark.parametrize("kind,N,nr,jit", [("lap7", 22, 4, "0"), ("lap7", 24, 2, "1"), ("lap7", 24, 2, "0"), ("lap27", 18, 3, "1"), ("elas", 8, 3, "0"), ("lap7", 24, 3, "1")])
def test_sharded_schedule_with_cyclic_supernode_single_process(kind, N, nr, jit):
    """Ranks emulated in one process on a mesh whose root supernode is wide enough (>= 512 columns) to be shared
    panel-cyclically ((24, 2): four panels on two ranks, so the just-in-time descendant updates of a rank's NEXT panel are
    exercised); checks that exactly all of L is broadcast once and that every rank ends with the oracle's factor."""
    import emulate_plan as E
    from suitesparse_b200 import gen
    from oracle import oracle
    # symbolic structure without the reference library: nested dissection of a 3-D mesh, supernodes = separators
    # (use the golden-free path: build the structure with the reference when present, else skip)
    from conftest import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build (host libcholmod for cholmod_l_analyze) not present")
    from suitesparse_b200.cholmod_host import Cholmod, _np_view
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem(kind, N)
    S = ch.sparse(A, +1); L = ch.analyze(S, p)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ch.factor_arrays(L).items()}
    S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
    Ap = _np_view(s2.p, n + 1, np.int64).copy(); Ai = _np_view(s2.i, int(Ap[n]), np.int64).copy(); Ax = _np_view(s2.x, int(Ap[n]), np.float64).copy()
    Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    st, minor, Lo = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], Sl)
    os.environ["SSB200_DIST_TAU"] = "0"              # cost model: free panel steps -> dominant supernodes are shared
    os.environ["SSB200_DIST_JIT"] = jit              # descendant updates just in time inside the panel loop, or up front
    try:
        plans = [E.export_plan(n, f["super"], f["pi"], f["px"], f["s"], nr, r) for r in range(nr)]
    finally:
        del os.environ["SSB200_DIST_TAU"], os.environ["SSB200_DIST_JIT"]
    if kind == "lap7":
        assert (plans[0]["owner"] < 0).sum() >= 1                  # a panel-cyclic supernode exists
    assert all(np.array_equal(pl["owner"], plans[0]["owner"]) for pl in plans)
    rel = E.relmap_of(plans[0], f["pi"], f["s"])
    Lx = [np.zeros(int(f["px"][-1])) for _ in range(nr)]
    for r in range(nr):
        E.assemble(plans[r], f["super"], f["pi"], f["px"], f["s"], Sl, Lx[r])
    nbytes = E.run_lockstep(plans, rel, Lx)
    assert nbytes == int(f["px"][-1])                              # all of L replicated exactly once
    for r in range(nr):
        assert persuper_relerr(f["px"], Lx[r], Lo) < 1e-11
    ch.free_sparse(S2); ch.free_factor(L)


@pytest.mark.parametrize("selective", [False, True])
@pytest.mark.parametrize("kind,N,nr", [("lap7", 22, 4), ("lap7", 24, 2), ("lap27", 18, 3), ("elas", 8, 3), ("lap7", 26, 8), ("lap7", 34, 3), ("lap7", 30, 2), ("lap7", 31, 3)])
def test_distributed_storage_schedule_single_process(kind, N, nr, selective):
    """The in-process multi-GPU path (ssb200_mg_*): every rank stores only its supernodes, the cyclic ones and the remote
    supernodes its updates read; finished ranges are pulled piecewise by the ranks that read them.  Emulated ranks with
    asynchronous pulls must reproduce the oracle's factor in the gathered host copy; local storage is smaller than L;
    every solve block belongs to exactly one rank."""
    import emulate_plan as E
    from suitesparse_b200 import gen
    from oracle import oracle
    from conftest import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build (host libcholmod for cholmod_l_analyze) not present")
    from suitesparse_b200.cholmod_host import Cholmod, _np_view
    ch = Cholmod(gpu=False)
    A, p = gen.make_problem(kind, N)
    S = ch.sparse(A, +1); L = ch.analyze(S, p)
    f = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ch.factor_arrays(L).items()}
    S2 = ch.lower_permuted(S, L); s2 = S2.contents; n = s2.nrow
    Ap = _np_view(s2.p, n + 1, np.int64).copy(); Ai = _np_view(s2.i, int(Ap[n]), np.int64).copy(); Ax = _np_view(s2.x, int(Ap[n]), np.float64).copy()
    Sl = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
    st, minor, Lo = oracle.factorize(n, f["super"], f["pi"], f["px"], f["s"], Sl)
    os.environ["SSB200_DIST_TAU"] = "0"
    if N in (30, 31):
        # the root supernode (4+ panels) is stored transiently: own panels packed, received panels in a ring of two slots
        os.environ["SSB200_MG_TRANSIENT_MIN"] = "1"; os.environ["SSB200_MG_RING"] = "2"
    if N == 34:
        # three supernodes of 255-289 columns on one level become panel-cyclic too: their chains are interleaved and start on
        # different ranks
        os.environ["SSB200_DIST_CYC_MIN"] = "200"
    try:
        plans = [E.export_plan_compact(n, f["super"], f["pi"], f["px"], f["s"], nr, r) for r in range(nr)]
    finally:
        del os.environ["SSB200_DIST_TAU"]
        os.environ.pop("SSB200_DIST_CYC_MIN", None); os.environ.pop("SSB200_MG_TRANSIENT_MIN", None); os.environ.pop("SSB200_MG_RING", None)
    if N in (30, 31):
        assert all(pl["tr"][:, 0].sum() == 1 for pl in plans)                  # the root is transient on every rank
        root = int(np.nonzero(plans[0]["tr"][:, 0])[0][0])
        full = int(f["px"][root + 1] - f["px"][root])
        nsrow_root = int(f["pi"][root + 1] - f["pi"][root])
        # nobody stores the whole root: own panels + two ring slots
        for pl in plans:
            nxt_base = pl["lxsize"] if root == len(f["px"]) - 2 else None
            assert pl["tr"][root, 2] + 2 * 256 * nsrow_root - pl["tr"][root, 1] < full
    xsize = int(f["px"][-1])
    if N == 34:
        # two panel-cyclic supernodes on one level (the children of the root separator): their chains are interleaved and
        # start on different ranks
        cyc = np.nonzero(plans[0]["owner"] < 0)[0]
        assert len(cyc) >= 4 and len(set(plans[0]["owner"][cyc].tolist())) >= 2, plans[0]["owner"][cyc]
    rel = E.relmap_of(plans[0], f["pi"], f["s"])
    Lx = [np.zeros(max(pl["lxsize"], 1)) for pl in plans]
    for r in range(nr):
        E.assemble_compact(plans[r], f["super"], f["pi"], f["px"], f["s"], Sl, Lx[r])
    pulled = E.run_lockstep_compact(plans, rel, Lx, f["px"], selective=selective)   # selective: a pull is delivered only when a step depends on it
    host = E.gather_compact(plans, Lx, f["px"], xsize)
    assert persuper_relerr(f["px"], host, Lo) < 1e-11
    # distributed: the ranks together pull less than a full replication would move, nobody stores everything (nr > 2)
    assert sum(pulled) < (nr - 1) * xsize
    if nr > 2:
        assert max(pl["lxsize"] for pl in plans) < xsize
    # solve blocks: each 64-column block of each supernode on exactly one rank, at a valid local offset
    nscol = np.diff(f["super"])
    blocks = int(np.ceil(nscol / 64).sum())
    cols = np.concatenate([pl["solve"][:, 2] for pl in plans])
    assert len(cols) == blocks and len(set(cols.tolist())) == blocks
    for pl in plans:
        assert (pl["solve"][:, 0] >= 0).all() and (pl["solve"][:, 0] < max(pl["lxsize"], 1)).all()
    ch.free_sparse(S2); ch.free_factor(L)
