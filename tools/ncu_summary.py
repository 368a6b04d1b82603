"""Fold `ncu --set full` reports of the level kernels into profiles/level_kernel_summary.json.

    python tools/ncu_summary.py KEY=report.ncu-rep[:cells_per_launch] ...

KEY is f64 / f32 (exhaustive kernel) or screen_f64 / screen_f32 (screened kernel); bench.py reads
dram_bytes_per_launch (roofline.traffic), warp_instructions_per_cell and issue_active_pct from this file."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "level_kernel_summary.json")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    try:
        doc = json.load(open(OUT))
    except Exception:
        doc = {}
    for spec in sys.argv[1:]:
        key, rest = spec.split("=", 1)
        rep, _, cells = rest.partition(":")
        cells = float(cells) if cells else 4999850001.0  # level j = 2 of C3
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        m = {h: {"value": float(v.replace(",", "")), "unit": u} for h, u, v in zip(hdr, units, vals) if h in KEEP}
        kname = vals[hdr.index("Kernel Name")]
        dram = sum(m[k]["value"] * UNIT.get(m[k]["unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        doc[key] = {"kernel": kname, "report": os.path.basename(rep), "cells_per_launch": cells,
                    "dram_bytes_per_launch": dram,
                    "warp_instructions_per_cell": m["smsp__inst_executed.sum"]["value"] / (cells / 32.0),
                    "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"],
                    "metrics": m}
    json.dump(doc, open(OUT, "w"), indent=1)
    for k, v in doc.items():
        if not isinstance(v, dict):
            continue
        print(k, v["kernel"][:60], "dram/launch %.2f MB" % (v["dram_bytes_per_launch"] / 1e6),
              "instr/warp-cell %.2f" % v["warp_instructions_per_cell"], "time", v["metrics"]["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
