"""Tier statistics of the screened kernel on BASELINE configuration 4 (null-model replicas, n=5000, T=8, f32): audit run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

ctx = capi.Context(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for rp in (False, True):
    a, b = synth.poisson_null(5000, seed=synth.SEED + 4, dtype=np.float32, replicas=R)
    ctx.solve_raw(a, b, 8, 1, rp, dtype=np.float32, want_perm=False, screen=2)
    t = ctx.timings()
    tk, td, sk, sd, gk, gd = t["screen_tiers"]
    cells = sum(t["level_cells"][:-1])
    print("rp", int(rp), "viol", t["screen_violations"], "tiles kept %d dropped %d (%.1f%%)" % (tk, td, 100 * tk / max(1, tk + td)),
          "| stages kept %.1f%% of all cells" % (100 * sk * 128 * 32 / cells),
          "| groups kept %.2f%%" % (100 * gk * 8 * 32 / cells), "| exact %.2f%%" % (100 * t["screen_exact_cells"] / cells))
    ctx.solve_raw(a, b, 8, 1, rp, dtype=np.float32, want_perm=False)
    t = ctx.timings()
    print("   levels", [round(x, 3) for x in t["level_ms"]], "total %.2f" % t["total_ms"])
