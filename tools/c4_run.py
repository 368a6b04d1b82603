"""BASELINE configuration 4 (4096 null-model replicas, n=5000, T=8, f32) as one batched host-buffer call."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
a, b = synth.poisson_null(5000, seed=synth.SEED + 4, dtype=np.float32, replicas=R)
for it in range(3):
    t0 = time.perf_counter()
    ctx.solve_raw(a, b, 8, 1, False, dtype=np.float32, want_perm=False, screen=scr, chunk_len=chunk)
    t = ctx.timings()
    print("R", R, "chunk", chunk, "screen", t["screen_used"], "wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(v, 2) for k, v in t.items() if k.endswith("_ms") and not isinstance(v, list)},
          "levels", [round(x, 2) for x in t["level_ms"]], "exact frac %.4f" % (t["screen_exact_cells"] / max(1, sum(t["level_cells"][:-1]))), flush=True)
