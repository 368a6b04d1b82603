"""C1 / C2 latency probe: wall time per call from host buffers, the engine's own stage times, launches.
    python tools/small_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth
ctx = capi.Context(0)
for name in ("C1", "C2"):
    cfg = synth.CONFIGS[name]
    dt = np.float32 if name == "C1" else np.float64
    a, b = cfg["gen"](dt)
    kw = dict(dtype=dt)
    for it in range(20):
        ctx.solve_raw(a, b, cfg["T"], cfg["dist"], cfg["risk_part"], **kw)
    t0 = time.perf_counter()
    N = 200
    for it in range(N):
        ctx.solve_raw(a, b, cfg["T"], cfg["dist"], cfg["risk_part"], **kw)
    wall = (time.perf_counter() - t0) / N
    t = ctx.timings()
    print(name, "n", len(a), "T", cfg["T"], dt.__name__, "wall %.1f us" % (1e6 * wall),
          {k: round(1e3 * t[k], 1) for k in ("total_ms", "h2d_ms", "sort_scan_ms", "prepass_ms", "levels_ms", "backtrace_ms", "d2h_ms")},
          "launches", t["kernel_launches"], flush=True)
