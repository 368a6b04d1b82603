"""Latency of one host-buffer call at the reference's own sizes (BASELINE configs 1 and 2): device-event stages vs wall."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

ctx = capi.Context(0)
for nm, dt in (("C1", np.float32), ("C1", np.float64), ("C2", np.float64)):
    c = synth.CONFIGS[nm]
    a, b = c["gen"](dt)
    for _ in range(5):
        ctx.solve_raw(a, b, c["T"], c["dist"], c["risk_part"], dtype=dt)
    t0 = time.perf_counter()
    N = 200
    for _ in range(N):
        ctx.solve_raw(a, b, c["T"], c["dist"], c["risk_part"], dtype=dt)
    wall = 1e3 * (time.perf_counter() - t0) / N
    t = ctx.timings()
    print(nm, np.dtype(dt).name, "wall %.3f ms/call" % wall, {k: round(v, 4) for k, v in t.items() if k.endswith("_ms") and not isinstance(v, list)},
          "launches", t["kernel_launches"], "levels", [round(x, 4) for x in t["level_ms"]])
