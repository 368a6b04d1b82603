"""Where does screening start to pay?  single instances and small batches, screen on/off, best of 5 calls."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
for dt in (np.float32, np.float64):
    for n, T, R in ((600, 5, 1), (1000, 5, 1), (2089, 4, 1), (3000, 8, 1), (5000, 8, 1), (8000, 8, 1), (1000, 5, 16), (600, 5, 64)):
        if R == 1:
            a, b = synth.poisson_planted(n, T, seed=1, dtype=dt)
        else:
            a, b = synth.poisson_null(n, seed=1, dtype=dt, replicas=R)
        res = {}
        for scr in (1, 0):
            best = 1e9
            for _ in range(6):
                t0 = time.perf_counter()
                ctx.solve_raw(a, b, T, 1, False, dtype=dt, screen=scr, want_perm=False)
                best = min(best, time.perf_counter() - t0)
            res[scr] = (best * 1e3, ctx.timings()["total_ms"], ctx.timings()["screen_used"])
        print(dt.__name__, n, T, R, "exhaustive wall %.3f dev %.3f | screened wall %.3f dev %.3f used %d" % (res[1][0], res[1][1], res[0][0], res[0][1], res[0][2]), flush=True)
