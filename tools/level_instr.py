"""Executed warp instructions per level of the screened path on the bench workload -> profiles/level_instr_r2.json.

    ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file f64.csv \
        python bench.py --steps 1 --warmup 3 --no-extras --skip-peaks [--dtype f32]
    python tools/level_instr.py f64=f64.csv f32=f32.csv

bench.py multiplies these per-level constants (the workload is seeded, so they are exact for it) by the levels it times and
divides by its own CUDA-event time: instruction-issue utilisation measured in the run, not a stored ratio."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "level_instr_r2.json")
KERNELS = {"screen_lb_kernel": "lower_bounds", "screen_select_kernel": "select", "level_live_screen_kernel": "live_tiles",
           "screen_fold_kernel": "fold"}
try:
    doc = json.load(open(OUT))
except Exception:
    doc = {}
for spec in sys.argv[1:]:
    key, path = spec.split("=", 1)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].split("<")[0].replace("void ", "").replace("ps::", "").strip()
        if name not in KERNELS:
            continue
        rec = per.setdefault((int(r[iid]), KERNELS[name]), {})
        rec[r[im]] = float(r[iv].replace(",", ""))
    by = collections.defaultdict(list)
    for (lid, k), rec in per.items():
        by[k].append((rec.get("smsp__inst_executed.sum", 0.0), rec.get("gpu__time_duration.sum", 0.0)))
    nlev = 14  # C3: levels 2..15 run the screened kernels, level 16 the last-level kernel
    doc[key] = {"n": 100000, "T": 16, "levels": nlev, "source": os.path.basename(path),
                "per_level": {k: [x[0] for x in v[-nlev:]] for k, v in by.items()},
                "ncu_us_per_level": {k: [x[1] / 1e3 for x in v[-nlev:]] for k, v in by.items()}}
    tot = sum(sum(v) for v in doc[key]["per_level"].values())
    print(key, "warp instructions per solve %.4g" % tot, {k: "%.3g" % sum(v) for k, v in doc[key]["per_level"].items()})
json.dump(doc, open(OUT, "w"), indent=1)
