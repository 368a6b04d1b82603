"""Tier statistics of the screened kernel on BASELINE configuration 3 (audit run: slow, every cell is also evaluated)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
cfg = synth.CONFIGS["C3"]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for dt in (np.float64, np.float32):
    a, b = cfg["gen"](dt)
    ctx.solve_raw(a, b, T, 1, False, dtype=dt, screen=2)
    t = ctx.timings()
    tk, td, sk, sd, gk, gd = t["screen_tiers"]
    cells = sum(t["level_cells"][:-1])
    print(dt.__name__, "viol", t["screen_violations"], "tiles kept %d dropped %d (%.1f%% kept)" % (tk, td, 100 * tk / max(1, tk + td)),
          "| warp-stages kept %d dropped %d (%.1f%% of all stages kept)" % (sk, sd, 100 * sk * 128 * 32 / cells),
          "| warp-groups kept %d dropped %d (%.2f%% of all groups kept)" % (gk, gd, 100 * gk * 8 * 32 / cells),
          "| exact %.2f%%" % (100 * t["screen_exact_cells"] / cells))
