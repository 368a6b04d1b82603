"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libps_ref.so).

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tools/make_golden.py

Each fixture stores the inputs, the reference's private DP state (priority_sortind_, every maxScore_
and nextStart_ row) and its public outputs (subsets, score_by_subset, get_optimal_score_extern(), and
the sweep table where asked).  The reference itself cannot travel to the GPU box, these files can.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_py as O  # noqa: E402
from partitionsolvers_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# reference golden vector, src/cpp/test/gtest_all.cpp:375-398 (inputs) -- expected subsets are
# asserted in tests/test_oracle.py from the same file's lines 392-398.
TB_A = [0.0212651, -0.20654906, -0.20654906, -0.20654906, -0.20654906, 0.0212651, -0.20654906, 0.0212651,
        -0.20654906, 0.0212651, -0.20654906, 0.0212651, -0.20654906, -0.06581402, 0.0212651, 0.03953075,
        -0.20654906, 0.16200014, 0.0212651, -0.20654906, 0.20296943, -0.18828341, -0.20654906, -0.20654906,
        -0.06581402, -0.20654906, 0.16200014, 0.03953075, -0.20654906, -0.20654906, 0.03953075, 0.20296943,
        -0.20654906, 0.0212651, 0.20296943, -0.20654906, 0.0212651, 0.03953075, -0.20654906, 0.03953075]
TB_B = [0.22771114, 0.21809504, 0.21809504, 0.21809504, 0.21809504, 0.22771114, 0.21809504, 0.22771114,
        0.21809504, 0.22771114, 0.21809504, 0.22771114, 0.21809504, 0.22682739, 0.22771114, 0.22745816,
        0.21809504, 0.2218354, 0.22771114, 0.21809504, 0.218429, 0.219738, 0.21809504, 0.21809504, 0.22682739,
        0.21809504, 0.2218354, 0.22745816, 0.21809504, 0.21809504, 0.22745816, 0.218429, 0.21809504, 0.22771114,
        0.218429, 0.21809504, 0.22771114, 0.22745816, 0.21809504, 0.22745816]
TB1_A = [2.26851454, 2.86139335, 5.51314769, 6.84829739, 6.96469186, 7.1946897, 9.80764198, 4.2310646]
TB1_B = [3.43178016, 3.92117518, 7.29049707, 7.37995406, 4.80931901, 4.38572245, 3.98044255, 0.59677897]


def cases():
    f32, f64 = np.float32, np.float64
    yield "testbaselines_gauss_rp_f32", dict(a=np.array(TB_A, f32), b=np.array(TB_B, f32), T=5, dist=0, rp=True,
                                             dtype=f32)
    yield "testbaselines_poisson_mc_f32", dict(a=np.array(TB1_A, f32), b=np.array(TB1_B, f32), T=3, dist=1,
                                               rp=False, dtype=f32)
    for dt, nm in ((f32, "f32"), (f64, "f64")):
        a, b = synth.CONFIGS["C1"]["gen"](dt)
        yield f"c1_poisson_mc_{nm}", dict(a=a, b=b, T=5, dist=1, rp=False, dtype=dt)
        yield f"c1_poisson_rp_{nm}", dict(a=a, b=b, T=5, dist=1, rp=True, dtype=dt)
        a, b = synth.CONFIGS["C2"]["gen"](dt)
        yield f"c2_gaussian_rp_{nm}", dict(a=a, b=b, T=4, dist=0, rp=True, dtype=dt)
        yield f"c2_gaussian_mc_{nm}", dict(a=a, b=b, T=4, dist=0, rp=False, dtype=dt, gamma=0.25, reg_power=2)
        a, b = synth.tract_like(2089, 4, seed=synth.SEED + 2, dtype=dt, dyadic_bits=10)
        yield f"c2_gaussian_rp_dyadic_{nm}", dict(a=a, b=b, T=4, dist=0, rp=True, dtype=dt)
        a, b = synth.poisson_planted(3000, 16, seed=synth.SEED + 3, dtype=dt)
        yield f"c3s_poisson_mc_{nm}", dict(a=a, b=b, T=16, dist=1, rp=False, dtype=dt)
        a, b = synth.gaussian_mixture(3000, 16, seed=synth.SEED + 3, dtype=dt)
        yield f"c3s_rational_{nm}", dict(a=a, b=b, T=16, dist=2, rp=False, dtype=dt)
        a, b = synth.gaussian_mixture(1500, 6, seed=synth.SEED + 7, dtype=dt, dyadic_bits=8)
        yield f"c3s_rational_dyadic_{nm}", dict(a=a, b=b, T=6, dist=2, rp=True, dtype=dt)
        a, b = synth.poisson_null(600, seed=synth.SEED + 4, dtype=dt, replicas=6)
        for r in range(6):
            yield f"c4s_null_r{r}_{nm}", dict(a=a[r], b=b[r], T=8, dist=1, rp=bool(r % 2), dtype=dt)
        a, b = synth.poisson_planted(700, 4, seed=synth.SEED + 8, dtype=dt)
        yield f"sweep_poisson_mc_{nm}", dict(a=a, b=b, T=12, dist=1, rp=False, dtype=dt, sweep=True)
        yield f"sweep_poisson_rp_{nm}", dict(a=a, b=b, T=12, dist=1, rp=True, dtype=dt, sweep=True, gamma=0.5,
                                             reg_power=1)
        a, b = synth.uniform_ab(400, seed=synth.SEED + 9, dtype=dt)
        yield f"timer_uniform_poisson_rp_{nm}", dict(a=a, b=b, T=10, dist=1, rp=True, dtype=dt)
        a, b = synth.uniform_ab(400, seed=synth.SEED + 10, dtype=dt, a_lo=-10.0)
        yield f"uniform_gaussian_mc_neg_{nm}", dict(a=a, b=b, T=6, dist=0, rp=False, dtype=dt)
        yield f"uniform_rational_neg_{nm}", dict(a=a, b=b, T=6, dist=2, rp=False, dtype=dt)
        a, b = synth.poisson_planted(64, 3, seed=synth.SEED + 11, dtype=dt)
        yield f"t_equals_n_{nm}", dict(a=a, b=b, T=64, dist=1, rp=False, dtype=dt)
        yield f"t_one_{nm}", dict(a=a, b=b, T=1, dist=1, rp=False, dtype=dt)
        yield f"t_over_n_{nm}", dict(a=a[:7], b=b[:7], T=20, dist=2, rp=True, dtype=dt)


def main():
    if not O.have("ref"):
        raise SystemExit("oracle/_ref/libps_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    total = 0
    for name, c in cases():
        sweep = c.get("sweep", False)
        r = O.solve("ref", c["a"], c["b"], c["T"], c["dist"], c["rp"], gamma=c.get("gamma", 0.0),
                    reg_power=c.get("reg_power", 1), sweep=sweep, dtype=c["dtype"])
        sizes = np.array([len(s) for s in r["subsets"]], np.int32)
        elems = np.concatenate(r["subsets"]).astype(np.int32) if len(r["subsets"]) else np.zeros(0, np.int32)
        blob = dict(a=c["a"], b=c["b"], T=np.int32(c["T"]), dist=np.int32(c["dist"]), rp=np.int32(c["rp"]),
                    gamma=np.float64(c.get("gamma", 0.0)), reg_power=np.int32(c.get("reg_power", 1)),
                    sweep=np.int32(sweep), perm=r["perm"], max_score=r["max_score"], next_start=r["next_start"],
                    subset_sizes=sizes, subset_elems=elems, score_by_subset=r["score_by_subset"],
                    score=r["score"], optimal_score_member=r["optimal_score_member"])
        if sweep:
            Tc = r["T"]
            all_sizes = np.zeros((Tc + 1, Tc), np.int32)
            all_elems = np.zeros((Tc + 1, r["n"]), np.int32)
            all_scores = np.zeros(Tc + 1, c["dtype"])
            for S, (ss, sc) in enumerate(r["all"]):
                all_scores[S] = sc
                off = 0
                for q, s in enumerate(ss):
                    all_sizes[S, q] = len(s)
                    all_elems[S, off:off + len(s)] = s
                    off += len(s)
            blob.update(all_sizes=all_sizes, all_elems=all_elems, all_scores=all_scores)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        total += os.path.getsize(path)
        print(f"{name:40s} n={len(c['a']):5d} T={c['T']:3d} score={float(r['score']):.9g}")
    print(f"wrote {total / 1024:.0f} KiB under {OUT}")


if __name__ == "__main__":
    main()
