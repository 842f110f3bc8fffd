import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
from suitesparse_b200.cholmod_host import Cholmod
from suitesparse_b200 import plain
import bench
ch = Cholmod(gpu=True)
A, perm, S, Lp, S2, t_an = bench.build_problem(ch, "lap7", 128)
f = ch.hot("cholmod_l_super_numeric"); beta = (C.c_double * 2)(0.0, 0.0)
f(S2, None, beta, Lp, C.byref(ch.cm)); pl = plain.plan_of_factor(Lp)
for it in range(3):
    t0 = time.perf_counter(); f(S2, None, beta, Lp, C.byref(ch.cm)); t1 = time.perf_counter()
    st = pl.stats()
    print("e2e wall %.1f ms | device total %.1f asm %.1f upd %.1f fac %.1f | h2d %.1f d2h_tail %.2f" % ((t1 - t0) * 1e3, st["ms_total"], st["ms_assemble"], st["ms_update"], st["ms_factor"], st["ms_h2d"], st["ms_d2h"]), "kinds", [round(v, 1) for v in st["ms_kind"][:5]])
for it in range(2):
    t0 = time.perf_counter(); pl.factorize_resident(); t1 = time.perf_counter(); st = pl.stats()
    print("resident wall %.1f ms | device total %.1f asm %.1f upd %.1f fac %.1f" % ((t1 - t0) * 1e3, st["ms_total"], st["ms_assemble"], st["ms_update"], st["ms_factor"]), "kinds", [round(v, 1) for v in st["ms_kind"][:5]])
import os
os.environ["SSB200_STREAM_D2H"] = "0"
