"""Cold start of the drop-in path: analyze, then the FIRST cholmod_l_super_numeric (plan build, L->x allocation + page-lock,
factorization, host copy) and the second one, timed on the host."""
import sys, time, ctypes as C, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from suitesparse_b200.cholmod_host import Cholmod
ch = Cholmod(gpu=True)
A, perm, S, Lp, S2, t_an = bench.build_problem(ch, "lap7", int(sys.argv[1]) if len(sys.argv) > 1 else 128)
f = ch.hot("cholmod_l_super_numeric"); beta = (C.c_double * 2)(0.0, 0.0)
for it in range(4):
    t0 = time.perf_counter(); ok = f(S2, None, beta, Lp, C.byref(ch.cm)); t1 = time.perf_counter()
    print("call %d: %.3f s  ok=%d status=%d" % (it, t1 - t0, ok, ch.cm.status), flush=True)
    if it == 0 and "--check" in sys.argv:
        # the host copy of the first call (staging ring + first-touch threads) against a plain download of the resident factor
        import numpy as np
        from suitesparse_b200 import plain
        pl = plain.plan_of_factor(Lp)
        ref = np.empty(Lp.contents.xsize, dtype=np.float64); pl.download_L(ref)
        x = ch.factor_arrays(Lp)["x"]
        bad = 0
        for a in range(0, x.size, 1 << 26):
            bad += int((x[a:a + (1 << 26)] != ref[a:a + (1 << 26)]).sum())
        print("first call: host L->x vs resident factor: %d of %d entries differ (staged=%d)" % (bad, x.size, pl.stats()["d2h_staged"]), flush=True)
        del ref
