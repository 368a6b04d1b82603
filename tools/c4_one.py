"""One C4-shaped batch (null-model replicas, n=5000, T=8, f32) for profiling: python tools/c4_one.py [replicas]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = capi.Context(0)
a, b = synth.poisson_null(5000, seed=synth.SEED + 4, dtype=np.float32, replicas=R)
for it in range(2):
    ctx.solve_raw(a, b, 8, 1, False, dtype=np.float32, want_perm=False)
t = ctx.timings()
print("total %.2f ms" % t["total_ms"], "levels", [round(x, 2) for x in t["level_ms"]])
