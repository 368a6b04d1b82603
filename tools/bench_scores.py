"""Time one C3-sized solve (n=100000, T=16) for every score x objective x precision (device events, best of 3).
    python tools/bench_scores.py  [n] [T]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = capi.Context()
names = {0: "Gaussian", 1: "Poisson", 2: "Rational"}
print(f"n={n} T={T}  (nominal cells = {synth.nominal_cells(n, T):.3e})")
for dt in (np.float64, np.float32):
    ap, bp = synth.poisson_planted(n, T, seed=synth.SEED + 3, dtype=dt)
    ag, bg = synth.gaussian_mixture(n, T, seed=synth.SEED + 3, dtype=dt)
    for dist in (1, 0, 2):
        for rp in (False, True):
            if dist == 2 and rp:
                continue
            a, b = (ap, bp) if dist == 1 else (ag, bg)
            best = 1e30
            for _ in range(3):
                ctx.solve_raw(a, b, T, dist, rp, dtype=dt, want_perm=False)
                best = min(best, ctx.timings()["total_ms"])
            print(f"{np.dtype(dt).name:8s} {names[dist]:9s} {'risk-part ' if rp else 'mult-clust'}  {best:8.1f} ms  "
                  f"{synth.nominal_cells(n, T) / (best * 1e-3):.3e} cells/s")
