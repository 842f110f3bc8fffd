"""Elimination-tree shard of one factorization over several B200s: one process per GPU (torchrun), NCCL over NVLink only
where the path has a real exchange step — the broadcast of a finished subtree / of a finished 256-column panel of a wide
top supernode to the ranks whose updates read it.  The schedule (who computes what, and which Lx range is
broadcast after which step) comes from the C++ plan builder (ssb_plan.cpp); this module only walks it:

    for every step k:   enqueue this rank's kernels of the step;   if the step ends with a broadcast: dist.broadcast(Lx[off:off+cnt], src)

Kernels and collectives share one CUDA stream, so no host synchronisation happens inside a factorization.
"""
from __future__ import annotations
import ctypes as C
import os
import numpy as np
from . import plain

c_long = C.c_int64


class _DevArray:
    """__cuda_array_interface__ view of raw device memory, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class ShardedFactor:
    def __init__(self, n, super_, pi, px, s, device_index: int, rank: int | None = None, world: int | None = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.lib = plain._lib()
        L = self.lib
        L.ssb200_plan_create_dist.restype = C.c_void_p
        L.ssb200_plan_create_dist.argtypes = [c_long, c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ssb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_dist_num_steps.restype = c_long; L.ssb200_dist_num_steps.argtypes = [C.c_void_p]
        L.ssb200_dist_step_info.argtypes = [C.c_void_p, c_long, C.POINTER(C.c_int), C.POINTER(c_long), C.POINTER(c_long), C.POINTER(C.c_int)]
        L.ssb200_dist_begin.argtypes = [C.c_void_p, C.c_void_p]
        L.ssb200_dist_run_step.argtypes = [C.c_void_p, c_long, C.c_int]
        L.ssb200_dist_end.argtypes = [C.c_void_p, C.POINTER(c_long)]
        L.ssb200_dist_zero_from.argtypes = [C.c_void_p, c_long]
        L.ssb200_dist_not_posdef.argtypes = [C.c_void_p, c_long, C.c_int, C.POINTER(C.c_int), C.POINTER(c_long), C.POINTER(c_long)]
        L.ssb200_dist_flops.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        a = [np.ascontiguousarray(v, dtype=np.int64) for v in (super_, pi, px, s)]
        self.n = int(n)
        h = L.ssb200_plan_create_dist(self.n, a[0].size - 1, *[v.ctypes.data_as(C.c_void_p) for v in a], device_index, self.world, self.rank)
        if not h:
            raise plain.SsbError(L.ssb200_last_error().decode())
        self.plan = plain.Plan(self.n, None, None, None, None, handle=h)
        self.plan.owned = True
        self.device = torch.device("cuda", device_index)
        self.stream = torch.cuda.Stream(device=self.device)          # kernels and NCCL broadcasts share this stream
        L.ssb200_set_stream(self.plan.h, C.c_void_p(self.stream.cuda_stream))
        self.xsize = self.plan.xsize
        self._copy_stream = None
        self.Lx = torch.as_tensor(_DevArray(L.ssb200_device_Lx(self.plan.h), max(self.xsize, 1)), device=self.device)
        ns = L.ssb200_dist_num_steps(self.plan.h)
        self.steps = []
        src, off, cnt, wt = C.c_int(), c_long(), c_long(), C.c_int()
        for k in range(ns):
            L.ssb200_dist_step_info(self.plan.h, k, C.byref(src), C.byref(off), C.byref(cnt), C.byref(wt))
            self.steps.append((src.value, off.value, cnt.value, wt.value))
        self.comm_stream = torch.cuda.Stream(device=self.device)     # broadcasts are issued here so they overlap the look-ahead work
        mine, total = C.c_double(), C.c_double()
        L.ssb200_dist_flops(self.plan.h, C.byref(mine), C.byref(total))
        self.my_flops, self.total_flops = mine.value, total.value

    def shared_host_factor(self, name: str = "ssb200_Lx"):
        """Host L->x in POSIX shared memory, mapped by every rank; each rank page-locks only the ranges it is the broadcast
        source of (its share of the factor) and copies exactly those out over its own PCIe link, instead of rank 0 pulling
        all of L.  Rank 0 is the application: after the factorization its view holds the whole factor.
        Returns a torch tensor of xsize doubles, or None (on every rank) if the shared buffer could not be set up."""
        torch, dist = self.torch, self.dist
        path = f"/dev/shm/{name}_{os.environ.get('MASTER_PORT', '0')}"
        nbytes = max(self.xsize, 1) * 8
        ok = 1
        t = None
        try:
            if self.rank == 0:
                with open(path, "wb") as f:
                    f.truncate(nbytes)
            if self.world > 1:
                dist.barrier()
            t = torch.from_file(path, shared=True, size=max(self.xsize, 1), dtype=torch.float64)
            # merge this rank's source ranges, round them to pages, pin them
            mine = sorted((off, off + cnt) for (src, off, cnt, _) in self.steps if src == self.rank and cnt > 0)
            merged = []
            for a, b in mine:
                a8, b8 = (a * 8) // 4096 * 4096, min(nbytes, -(-(b * 8) // 4096) * 4096)
                if merged and a8 <= merged[-1][1]:
                    merged[-1][1] = max(merged[-1][1], b8)
                else:
                    merged.append([a8, b8])
            base = t.data_ptr()
            self._pinned = []
            for a8, b8 in merged:
                rc = torch.cuda.cudart().cudaHostRegister(base + a8, b8 - a8, 0)
                if int(rc) != 0:
                    ok = 0
                    break
                self._pinned.append(base + a8)
        except Exception:
            ok = 0
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
        if self.world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if self.rank == 0 and os.path.exists(path):
            os.unlink(path)                               # the mappings keep it alive
        if not ok:
            for ptr in getattr(self, "_pinned", []):
                torch.cuda.cudart().cudaHostUnregister(ptr)
            self._pinned = []
            return None
        self._shared = t
        return t

    def upload_A(self, A_lower, F=None):
        return self.plan.upload_A(A_lower, F)

    def factorize_resident(self, beta: float = 0.0, host_out=None, host_shared: bool = False, quick_return: bool = False):
        """Returns (status, minor): status 0 ok, 1 not positive definite (every rank gets the same answer).
        host_out: pinned host tensor of xsize doubles (or None).  A range of L is final on every rank right after its
        broadcast, so it is copied to host_out on a second stream while the factorization continues.
        host_shared: host_out is the same shared-memory buffer on every rank (shared_host_factor()): each rank copies
        only the ranges it is the source of, as soon as it has finished them."""
        torch, dist, L, h = self.torch, self.dist, self.lib, self.plan.h
        b = (C.c_double * 2)(beta, 0.0)
        if host_out is not None and self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        import os
        dbg = os.environ.get("SSB200_DIST_DEBUG", "")      # "nocomm" / "nocompute": timing experiments only (wrong factor)
        do_comm, do_compute = dbg != "nocomm", dbg != "nocompute"
        self.phase_event = None
        self.step_events = [] if dbg or os.environ.get("SSB200_DIST_STEPTIME") else None
        with torch.cuda.stream(self.stream):
            self.plan._check(L.ssb200_dist_begin(h, b))
            pending = []                                  # collectives in flight: work handles
            deferred = []                                 # finished subtrees whose exchange has not been issued yet
            batch_subtrees = (self.world > 1 and do_comm and os.environ.get("SSB200_DIST_P2P", "0") != "0"
                              and (host_out is None or host_shared))

            def flush_deferred():
                # Optional (SSB200_DIST_P2P=1): all finished subtrees change hands in ONE grouped exchange of point-to-point
                # sends instead of one broadcast per subtree.  Measured slower on 8 B200 (308 vs 296 ms at lap7 128^3), so
                # the default stays with asynchronous broadcasts.
                if not deferred:
                    return
                self.comm_stream.wait_stream(self.stream)
                with torch.cuda.stream(self.comm_stream):
                    ops = []
                    for (src, off, cnt) in deferred:
                        view = self.Lx[off:off + cnt]
                        if src == self.rank:
                            ops += [dist.P2POp(dist.isend, view, peer) for peer in range(self.world) if peer != self.rank]
                        else:
                            ops.append(dist.P2POp(dist.irecv, view, src))
                    pending.extend(dist.batch_isend_irecv(ops))
                deferred.clear()

            for k, (src, off, cnt, wait_remote) in enumerate(self.steps):
                if self.step_events is not None:
                    ev = torch.cuda.Event(enable_timing=True); ev.record(); self.step_events.append(ev)
                if wait_remote and self.phase_event is None:
                    self.phase_event = torch.cuda.Event(enable_timing=True); self.phase_event.record()   # subtree phase ends here
                if wait_remote:
                    flush_deferred()
                if wait_remote and pending:
                    for w in pending:
                        w.wait()                          # stream-level: self.stream waits for the NCCL stream
                    pending = []
                if do_compute:
                    self.plan._check(L.ssb200_dist_run_step(h, k, 0))
                w = None
                if src >= 0:
                    if batch_subtrees and not wait_remote:
                        deferred.append((src, off, cnt))  # subtree phase: nobody reads remote data yet
                    elif self.world > 1 and do_comm:
                        self.comm_stream.wait_stream(self.stream)
                        with torch.cuda.stream(self.comm_stream):
                            w = dist.broadcast(self.Lx[off:off + cnt], src, async_op=True)
                        pending.append(w)
                    if host_out is not None and host_shared:
                        if src == self.rank:              # final here since the launches above: no need to wait for the transfer
                            self._copy_stream.wait_stream(self.stream)
                            with torch.cuda.stream(self._copy_stream):
                                host_out[off:off + cnt].copy_(self.Lx[off:off + cnt], non_blocking=True)
                    elif host_out is not None:
                        # rank 0 pulls everything: the range is final there once its broadcast has landed
                        with torch.cuda.stream(self._copy_stream):
                            if w is not None:
                                w.wait()
                            else:
                                self._copy_stream.wait_stream(self.stream)
                            host_out[off:off + cnt].copy_(self.Lx[off:off + cnt], non_blocking=True)
                if do_compute:
                    self.plan._check(L.ssb200_dist_run_step(h, k, 1))
            flush_deferred()
            for w in pending:
                w.wait()
            if self.step_events is not None:
                ev = torch.cuda.Event(enable_timing=True); ev.record(); self.step_events.append(ev)
            bad = c_long(self.n)
            self.plan._check(L.ssb200_dist_end(h, C.byref(bad)))
            minor = bad.value
            if self.world > 1:
                t = torch.tensor([minor], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                minor = int(t.item())
            if minor < self.n:
                # the reference's protocol (t_cholmod_super_numeric.c:905-968): zero from the failing supernode on, then one
                # rank repeats that supernode on the columns before the failing one and hands the block to the others
                redo, roff, rcnt = C.c_int(-1), c_long(0), c_long(0)
                self.plan._check(L.ssb200_dist_not_posdef(h, minor, 1 if quick_return else 0, C.byref(redo), C.byref(roff), C.byref(rcnt)))
                if redo.value >= 0 and self.world > 1:
                    dist.broadcast(self.Lx[roff.value:roff.value + rcnt.value], redo.value)
                    torch.cuda.current_stream().synchronize()
                if host_out is not None:
                    self._copy_stream.synchronize()
                    if not host_shared or self.rank == 0:
                        host_out[:self.xsize].copy_(self.Lx[:self.xsize])
                    if host_shared and self.world > 1:
                        dist.barrier()
                return 1, minor
        if host_out is not None:
            if self.world == 1:                           # nothing is broadcast with one rank: plain copy at the end
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_stream(self.stream)
                    host_out[:self.xsize].copy_(self.Lx[:self.xsize], non_blocking=True)
            self._copy_stream.synchronize()
            if host_shared and self.world > 1:
                dist.barrier()                            # every rank's share has landed in the shared buffer
        return 0, self.n

    def download_L(self, out=None):
        self.stream.synchronize()
        return self.plan.download_L(out)

    def solve(self, X, which: int = 2):
        """Every rank holds the complete factor after a factorization, so the solve is local (replicated)."""
        self.stream.synchronize()
        return self.plan.solve(X, which)

    def close(self):
        self.plan.close()
