"""C3 timing (f64 and f32) across chunk lengths + C4 replicas: quick probe of the screened path.

    python tools/c3_probe.py [chunk,chunk,...]
    PS_DEBUG_LEVEL=5 PS_DEBUG_FILE=gpurun_out/tile_times.bin python tools/c3_probe.py 0   # then tools/tile_times.py
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
cfg = synth.CONFIGS["C3"]
for dt in (np.float64, np.float32):
    a, b = cfg["gen"](dt)
    for chunk in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,512,1024,2048,4096").split(",")]:
        best = 1e9
        for it in range(4):
            ctx.solve_raw(a, b, 16, 1, False, dtype=dt, want_perm=False, chunk_len=chunk)
            t = ctx.timings()
            best = min(best, t["total_ms"])
        print(dt.__name__, "chunk", chunk, "best total %.3f ms" % best, "levels %.3f" % t["levels_ms"],
              "lev", [round(x, 3) for x in t["level_ms"]], "exact frac %.4f" % (t["screen_exact_cells"] / sum(t["level_cells"][:-1])), flush=True)
a, b = synth.poisson_null(5000, seed=synth.SEED + 4, dtype=np.float32, replicas=4096)
for chunk in (0,):
    for it in range(2):
        t0 = time.perf_counter()
        ctx.solve_raw(a, b, 8, 1, False, dtype=np.float32, want_perm=False, chunk_len=chunk)
        t = ctx.timings()
    print("C4 chunk", chunk, "wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), "total %.2f" % t["total_ms"], "levels", [round(x, 2) for x in t["level_ms"]],
          "exact frac %.4f" % (t["screen_exact_cells"] / sum(t["level_cells"][:-1])), flush=True)
