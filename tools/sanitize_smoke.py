"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the library on small inputs.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402

ctx = capi.Context()
for dt in (np.float32, np.float64):
    a, b = synth.poisson_planted(1537, 5, seed=3, dtype=dt)
    for dist, rp in ((1, False), (1, True), (0, False), (2, False)):
        for sm in (capi.PS_SUM_RUNNING, capi.PS_SUM_PREFIX):
            ctx.solve_raw(a, b, 6, dist, rp, dtype=dt, sum_mode=sm, sweep=True, levels=True, chunk_len=256)
            ctx.solve_raw(a, b, 6, dist, rp, dtype=dt, sum_mode=sm, no_fast=1)
    ab, bb = synth.poisson_null(301, seed=4, dtype=dt, replicas=5)
    ctx.solve_raw(ab, bb, 4, 1, False, dtype=dt, sweep=True)
    ctx.solve_raw(a[:1], b[:1], 1, 1, False, dtype=dt)
    ctx.solve_raw(a[:129], b[:129], 129, 2, False, dtype=dt, levels=True)
    ctx.range_score(a, b, 1, False, dtype=dt)
    ctx.eval_log(np.abs(a) + 0.1, dtype=dt)
    # screened levels (csrc/screen.cuh): every score, plain and audit mode, a replica batch, an unsorted permutation
    a6, b6 = synth.poisson_planted(6143, 4, seed=6, dtype=dt)
    for dist, rp in ((1, False), (1, True), (0, False), (0, True), (2, False)):
        ctx.solve_raw(a6, b6, 4, dist, rp, dtype=dt, levels=True)
        assert ctx.timings()["screen_used"] == 1
    ctx.solve_raw(a6, b6, 4, 1, False, dtype=dt, screen=2)
    ctx.solve_raw(a6, b6, 4, 1, False, dtype=dt, perm=np.random.RandomState(1).permutation(6143).astype(np.int32))
    ab, bb = synth.poisson_null(1531, seed=4, dtype=dt, replicas=5)
    ctx.solve_raw(ab, bb, 4, 1, False, dtype=dt, sweep=True)
    assert ctx.timings()["screen_used"] == 1
# running-sum screens (csrc/screen_run.cuh): real-valued inputs, some a <= 0; audit mode takes them at any size
for dt in (np.float64, np.float32):
    ar, br = synth.gaussian_mixture(6143, 6, seed=7, dtype=dt)
    for dist, rp in ((1, False), (1, True), (0, False), (0, True), (2, False)):
        aa = np.maximum(ar, 0.5) if dist == 1 else ar
        ctx.solve_raw(aa, br, 4, dist, rp, dtype=dt, levels=True, sweep=True, screen=2)
        t = ctx.timings()
        assert t["screen_violations"] == 0
        assert t["screen_used"] == (2 if (dt == np.float64 or dist != 1) else 0), (dt, dist, rp, t["screen_used"])
    ar, br = synth.gaussian_mixture(16511, 4, seed=8, dtype=dt)
    ctx.solve_raw(ar, br, 3, 2, False, dtype=dt)
    assert ctx.timings()["screen_used"] == 2
    ctx.solve_raw(ar, br, 3, 2, False, dtype=dt, perm=np.random.RandomState(2).permutation(16511).astype(np.int32))
ctx.selftest_mufu()
ctx.ltss(a.astype(np.float32), b.astype(np.float32), 1)
ctx.selftest_math(0, 1 << 16)
ctx.selftest_math(3, 1 << 16)
try:
    import torch
    from partitionsolvers_b200 import multigpu
    dev = torch.device("cuda", ctx.device)
    a, b = synth.poisson_planted(6200, 4, seed=5, dtype=np.float64)
    engines = [multigpu.CudaShardEngine(ctx, dev) for _ in range(2)]
    perm, a_s, b_s = engines[0].sort(a, b, np.float64)
    for r, e in enumerate(engines):
        e.create(a_s, b_s, 6200, 4, 1, False, np.float64, r, 2)
    chunks = [e.new_chunk() for e in engines]
    for e, c in zip(engines, chunks):
        e.level1(c)
    g = torch.cat(chunks)
    for j in range(2, 5):
        for e, c in zip(engines, chunks):
            e.set_prev(g)
            e.level(j, c)
        g = torch.cat(chunks)
    for e in engines:
        e.close()
except ImportError:
    pass
ctx.close()
print("sanitize_smoke: done")
