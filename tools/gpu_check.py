"""First-contact / debugging run on a GPU box: CUDA path vs the CPU oracle, verbose.

    python tools/gpu_check.py [--big]

Prints, per case, how many maxScore_/nextStart_ cells differ from the oracle (port of the reference;
oracle/dp_oracle.cpp) and the worst relative error.  Not a test (tests/ holds those) -- a diagnostic.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oracle_py as O  # noqa: E402
from partitionsolvers_b200 import capi, host, synth  # noqa: E402


def compare(raw, ora, tag):
    ms_g, ms_o = raw["max_score"], ora["max_score"]
    ns_g, ns_o = raw["next_start"], ora["next_start"]
    same_perm = np.array_equal(raw["perm"], ora["perm"])
    neq_ms = int((ms_g != ms_o).sum())
    with np.errstate(all="ignore"):
        rel = np.abs(ms_g.astype(np.float64) - ms_o.astype(np.float64)) / np.maximum(
            np.abs(ms_o.astype(np.float64)), 1e-300)
    rel = np.where(ms_g == ms_o, 0.0, rel)
    neq_ns = int((ns_g != ns_o).sum())
    print(f"{tag:58s} perm={'ok' if same_perm else 'DIFF'} ms!= {neq_ms:7d} (max rel {rel.max():.2e}) "
          f"ns!= {neq_ns:6d} / {ns_o.size}")
    return same_perm, neq_ms, neq_ns, rel.max()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    O.build(("port",))
    ctx = capi.Context()
    print("device", ctx.device)
    cases = []
    for dt in (np.float64, np.float32):
        cases.append(("C1 poisson MC", synth.CONFIGS["C1"]["gen"](dt), 5, 1, False, dt))
        cases.append(("C1 poisson RP", synth.CONFIGS["C1"]["gen"](dt), 5, 1, True, dt))
        cases.append(("C2 gaussian RP", synth.CONFIGS["C2"]["gen"](dt), 4, 0, True, dt))
        cases.append(("C2 gaussian MC", synth.CONFIGS["C2"]["gen"](dt), 4, 0, False, dt))
        cases.append(("gmix rational n=3000 T=7", synth.gaussian_mixture(3000, 7, dtype=dt), 7, 2, False, dt))
        cases.append(("gmix dyadic rational n=3000 T=7", synth.gaussian_mixture(3000, 7, dtype=dt, dyadic_bits=6),
                      7, 2, False, dt))
        cases.append(("tiny n=5 T=3", synth.poisson_planted(5, 2, dtype=dt), 3, 1, False, dt))
        cases.append(("n=129 T=129", synth.poisson_planted(129, 3, dtype=dt), 129, 1, False, dt))
    for name, (a, b), T, dist, rp, dt in cases:
        ora = O.solve("port", a, b, T, dist, rp, dtype=dt)
        for sm, smn in ((capi.PS_SUM_RUNNING, "running"), (capi.PS_SUM_PREFIX, "prefix")):
            for chunk in (0, 256):
                raw = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, sum_mode=sm, perm=ora["perm"], levels=True,
                                    chunk_len=chunk)
                compare(raw, ora, f"{dt.__name__} {name} {smn} L={chunk}")
        # device sort
        raw = ctx.solve_raw(a, b, T, dist, rp, dtype=dt, levels=True)
        s = host.DPSolver(len(a), T, a, b, dist, rp, dtype=dt, ctx=ctx)
        print(f"   device-sort: perm equal={np.array_equal(raw['perm'], ora['perm'])} "
              f"score gpu={s.get_optimal_score_extern():.10g} oracle={ora['score']:.10g}")
    if args.big:
        for dt in (np.float32, np.float64):
            cfg = synth.CONFIGS["C3"]
            a, b = cfg["gen"](dt)
            for sm, smn in ((capi.PS_SUM_RUNNING, "running"), (capi.PS_SUM_PREFIX, "prefix")):
                for it in range(2):
                    t0 = time.time()
                    raw = ctx.solve_raw(a, b, cfg["T"], cfg["dist"], cfg["risk_part"], dtype=dt, sum_mode=sm)
                    t1 = time.time()
                tm = ctx.timings()
                cells = synth.nominal_cells(cfg["n"], cfg["T"])
                print(f"C3 {dt.__name__} {smn}: wall {t1 - t0:.3f}s  total_ms {tm['total_ms']:.1f} "
                      f"levels_ms {tm['levels_ms']:.1f} sort {tm['sort_scan_ms']:.2f} pre {tm['prepass_ms']:.2f} "
                      f"bt {tm['backtrace_ms']:.2f} -> {cells / (tm['total_ms'] * 1e-3):.3e} nominal cells/s")
                print("   level_ms", [round(x, 2) for x in tm["level_ms"]])
                print("   bounds", raw["bounds"][0].tolist())


if __name__ == "__main__":
    main()
