"""Development aid: summarise the per-tile timeline the library records when PS_DEBUG_LEVEL=j is set (start / end / SM / cost
class of every live tile of level j of the last solve; written to $PS_DEBUG_FILE or ./tile_times.bin)."""
import sys
import numpy as np
d = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tile_times.bin", dtype=np.uint64).reshape(-1, 4)
d = d[d[:, 0] > 0]
t0 = d[:, 0].astype(np.int64); t1 = d[:, 1].astype(np.int64)
cls = (d[:, 2] & np.uint64(0xffffffff)).astype(int)
base = t0.min(); dur = (t1 - t0) / 1e3
print("tiles", len(d), "span us %.0f" % ((t1.max() - base) / 1e3), "CTA-time ms %.1f -> %.0f us per slot (1184)" % (dur.sum() / 1e3, dur.sum() / 1184))
for k in range(4):
    m = cls == k
    if m.any():
        print("class", k, "n", m.sum(), "dur mean %.1f med %.1f p90 %.1f max %.1f sum(ms) %.2f" % (dur[m].mean(), np.median(dur[m]), np.percentile(dur[m], 90), dur[m].max(), dur[m].sum() / 1e3),
              "start us %.0f..%.0f" % ((t0[m].min() - base) / 1e3, (t0[m].max() - base) / 1e3))
edges = np.linspace(0, (t1.max() - base) / 1e3, 17)
print("busy CTAs per interval:", " ".join("%d" % (np.clip(np.minimum((t1 - base) / 1e3, hi) - np.maximum((t0 - base) / 1e3, lo), 0, None).sum() / (hi - lo)) for lo, hi in zip(edges[:-1], edges[1:])))
