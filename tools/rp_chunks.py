"""Poisson risk-partitioning on the C3 counts across chunk lengths.  python tools/rp_chunks.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
for dt in (np.float64, np.float32):
    a, b = synth.CONFIGS["C3"]["gen"](dt)
    for chunk in (0, 256, 1024):
        best = 1e9
        for it in range(3):
            ctx.solve_raw(a, b, 16, 1, True, dtype=dt, want_perm=False, chunk_len=chunk)
            t = ctx.timings()
            best = min(best, t["total_ms"])
        print(dt.__name__, "chunk", chunk, "best %.2f ms" % best, "exact frac %.4f" % (t["screen_exact_cells"] / sum(t["level_cells"][:-1])), flush=True)
