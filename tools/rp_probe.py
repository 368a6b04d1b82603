"""Poisson risk-partitioning at the C3 shape (counts f64/f32, real-valued f64/f32): time, exact fraction, audit.
    python tools/rp_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
for dt in (np.float64, np.float32):
    for kind in ("counts", "real"):
        if kind == "counts":
            a, b = synth.CONFIGS["C3"]["gen"](dt)
        else:
            a, b = synth.CONFIGS["C3R"]["gen"](dt)
            a = np.abs(a) + dt(0.25)
        best = 1e9
        for it in range(3):
            ctx.solve_raw(a, b, 16, 1, True, dtype=dt, want_perm=False)
            t = ctx.timings()
            best = min(best, t["total_ms"])
        frac = t["screen_exact_cells"] / sum(t["level_cells"][:-1])
        ctx.solve_raw(a, b, 4, 1, True, dtype=dt, want_perm=False, screen=2)
        ta = ctx.timings()
        tk, td, sk, sd, gk, gd = ta["screen_tiers"]
        print(dt.__name__, kind, "best %.2f ms" % best, "screen", t["screen_used"], "exact %.4f" % frac,
              "| audit T=4: viol", ta["screen_violations"], "tiles kept %.1f%%" % (100 * tk / max(1, tk + td)),
              "groups kept/dropped", gk, gd, flush=True)
