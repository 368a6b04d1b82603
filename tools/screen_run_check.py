"""GPU check of the running-sum screened kernel (screen_run.cuh) on real-valued double inputs: tables identical to the
exhaustive kernel (every maxScore_/nextStart_ cell), audit clean, timing.   python tools/screen_run_check.py [quick] [f32]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

quick = "quick" in sys.argv[1:]
ctx = capi.Context(0)
dt = np.float32 if "f32" in sys.argv[1:] else np.float64
SCORES = ((1, False), (1, True), (0, False), (0, True), (2, False))
if dt == np.float32:
    SCORES = ((0, False), (0, True), (2, False))  # the float running-sum screen covers Gaussian and Rational


def gens(n, T):
    yield "gauss_mix", synth.gaussian_mixture(n, T, seed=synth.SEED + 3, dtype=dt)
    yield "tract", synth.tract_like(n, max(2, T // 2), seed=synth.SEED + 2, dtype=dt)
    yield "uniform", synth.uniform_ab(n, seed=synth.SEED + 7, dtype=dt, a_lo=0.5, b_lo=0.5)
    a, b = synth.gaussian_mixture(n, T, seed=synth.SEED + 9, dtype=dt)
    yield "gauss_pos", (np.abs(a) + 0.25, b)


bad = 0
for n, T in ((1000, 5), (2089, 4), (6000, 8), (20000, 8)) + (() if quick else ((100000, 16),)):
    for name, (a, b) in gens(n, T):
        for dist, rp in SCORES:
            if dist == 1 and (a <= 0).any():
                a = np.maximum(a, 0.125)
            if n >= 100000 and name not in ("gauss_mix", "gauss_pos"):
                continue
            lv = n <= 20000
            r1 = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=lv, screen=1)
            t1 = ctx.timings()
            r0 = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=lv, screen=0)
            t0 = ctx.timings()
            ok = np.array_equal(r0["bounds"], r1["bounds"]) and np.array_equal(r0["subset_scores"], r1["subset_scores"])
            if lv:
                ok = ok and np.array_equal(r0["max_score"], r1["max_score"]) and \
                    np.array_equal(r0["next_start"], r1["next_start"])
            msg = ""
            if n <= 20000:
                ra = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True, screen=2)
                ta = ctx.timings()
                ok = ok and np.array_equal(ra["max_score"], r1["max_score"]) and \
                    np.array_equal(ra["next_start"], r1["next_start"])
                msg = "audit viol %d (used %d)" % (ta["screen_violations"], ta["screen_used"])
                if ta["screen_violations"]:
                    ok = False
            cells = sum(t0["level_cells"][:-1])
            bad += 0 if ok else 1
            print(name, n, T, "dist", dist, "rp", int(rp), "same" if ok else "DIFF", "used", t0["screen_used"],
                  "levels ms exh %.3f scr %.3f" % (t1["levels_ms"], t0["levels_ms"]),
                  "prepass %.3f/%.3f" % (t1["prepass_ms"], t0["prepass_ms"]),
                  "total %.3f -> %.3f" % (t1["total_ms"], t0["total_ms"]),
                  "exact frac %.4f" % (t0["screen_exact_cells"] / max(cells, 1)), msg, flush=True)
            if n >= 100000:
                print("  level ms", ["%.2f" % x for x in t0["level_ms"]])
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
