"""Quick GPU check of the screened level kernel: tables identical to the exhaustive kernel, audit clean, timing."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth

ctx = capi.Context(0)
print("mufu (units of 2^-24): rcp %.3f lg2 %.3f" % ctx.selftest_mufu())
cases = [("C1", 1000, 5), ("C3", 20000, 8), ("C3", 100000, 16), ("C4", 5000, 8)]
for name, n, T in cases:
    cfg = synth.CONFIGS[name]
    for dt in (np.float64, np.float32):
        if name == "C4":
            a, b = cfg["gen"](dt, n, 64)
        else:
            a, b = cfg["gen"](dt, n)
        for dist, rp in ((1, False), (1, True), (0, False), (0, True), (2, False)):
            if n >= 100000 and (dist, rp) != (1, False):
                continue
            lv = n <= 20000
            t0 = time.time(); r1 = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=lv, screen=1); t1 = ctx.timings()
            r0 = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=lv, screen=0); t0_ = ctx.timings()
            ok = np.array_equal(r0["bounds"], r1["bounds"]) and np.array_equal(r0["subset_scores"], r1["subset_scores"])
            if lv:
                ok = ok and np.array_equal(r0["max_score"], r1["max_score"]) and np.array_equal(r0["next_start"], r1["next_start"])
            msg = ""
            if n <= 20000:
                ctx.solve_raw(a, b, T, dist, rp, dtype=dt, screen=2); ta = ctx.timings()
                msg = "audit viol %d" % ta["screen_violations"]
            cells = sum(t0_["level_cells"][:-1])
            print(name, n, T, dt.__name__, "dist", dist, "rp", int(rp), "same" if ok else "DIFF", "used", t0_["screen_used"],
                  "levels ms exh %.3f scr %.3f" % (t1["levels_ms"], t0_["levels_ms"]),
                  "total %.3f -> %.3f" % (t1["total_ms"], t0_["total_ms"]),
                  "exact frac %.4f" % (t0_["screen_exact_cells"] / max(cells, 1)), msg, flush=True)
            if n >= 100000:
                print("  level ms", ["%.2f" % x for x in t0_["level_ms"]])
