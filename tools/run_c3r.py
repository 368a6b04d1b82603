"""C3R (n=100000, T=16 Rational score on real-valued double data): timing, tier statistics, activation threshold.
    python tools/run_c3r.py time|tiers|threshold [f32]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
ctx = capi.Context(0)
dt = np.float32 if "f32" in sys.argv[2:] else np.float64
if mode == "time":
    a, b = synth.CONFIGS["C3R"]["gen"](dt)
    for it in range(4):
        ctx.solve_raw(a, b, 16, 2, False, dtype=dt, want_perm=False)
        t = ctx.timings()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items() if not isinstance(v, list)})
    print("level ms", ["%.2f" % x for x in t["level_ms"]])
elif mode == "tiers":
    a, b = synth.CONFIGS["C3R"]["gen"](dt)
    for dist, rp in ((2, False), (0, False), (1, False)):
        aa = np.maximum(a, 0.125) if dist == 1 else a
        ctx.solve_raw(aa, b, 5, dist, rp, dtype=dt, screen=2)
        t = ctx.timings()
        tk, td, sk, sd, gk, gd = t["screen_tiers"]
        cells = sum(t["level_cells"][:-1])
        print("dist", dist, "used", t["screen_used"], "viol", t["screen_violations"],
              "tiles kept %.1f%%" % (100 * tk / max(1, tk + td)),
              "| stages kept %.1f%% of all" % (100 * sk * 128 * 32 / cells),
              "| groups kept %.2f%% of all" % (100 * gk * 8 * 32 / cells),
              "| exact %.2f%%" % (100 * t["screen_exact_cells"] / cells))
else:
    for n in (6000, 8000, 10000, 12000, 14000, 16000, 20000):
        a, b = synth.gaussian_mixture(n, 8, seed=synth.SEED + 3, dtype=dt)
        out = []
        for scr in (1, 0):
            best = 1e9
            for _ in range(5):
                ctx.solve_raw(a, b, 8, 2, False, dtype=dt, screen=scr, want_perm=False)
                best = min(best, ctx.timings()["total_ms"])
            out.append((best, ctx.timings()["screen_used"]))
        print(n, "exhaustive %.3f ms | default %.3f ms (used %d)" % (out[0][0], out[1][0], out[1][1]))
