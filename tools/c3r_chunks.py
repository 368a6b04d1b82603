"""C3R (real-valued, Rational) across chunk lengths: time and exactly evaluated fraction.  python tools/c3r_chunks.py [f32]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
dt = np.float32 if "f32" in sys.argv[1:] else np.float64
ctx = capi.Context(0)
a, b = synth.CONFIGS["C3R"]["gen"](dt)
for chunk in (0, 256, 512, 1024, 2048):
    best = 1e9
    for it in range(3):
        ctx.solve_raw(a, b, 16, 2, False, dtype=dt, want_perm=False, chunk_len=chunk)
        t = ctx.timings()
        best = min(best, t["total_ms"])
    print(dt.__name__, "chunk", chunk, "best %.2f ms" % best, "levels %.2f" % t["levels_ms"], "prepass %.2f" % t["prepass_ms"],
          "exact frac %.4f" % (t["screen_exact_cells"] / sum(t["level_cells"][:-1])), flush=True)
