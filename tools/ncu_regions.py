"""Summarise an ncu report of a level kernel: headline metrics + executed instructions / stall samples per code region
(regions = straight-line runs between branches).  usage: python tools/ncu_regions.py report.ncu-rep cells_per_launch"""
import csv, subprocess, sys, io
rep, cells = sys.argv[1], float(sys.argv[2]) / 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
for h, u, v in zip(hdr, rows[1], vals):
    if h in want:
        print("%-70s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
print("instructions per warp-cell %.2f, samples %d" % (tot / cells, tots))
start = 0
for idx, d in enumerate(data):
    if any(x in d[0] for x in ("BRA", "EXIT", "BSYNC", "VOTE")) or idx == len(data) - 1:
        seg = data[start:idx + 1]
        s, sm = sum(x[1] for x in seg), sum(x[2] for x in seg)
        if s / cells > 0.04 or sm / tots > 0.02:
            print("%4d-%4d  inst/cell %.3f  samples %4.1f%%   %s ... %s" % (start, idx, s / cells, 100 * sm / tots, seg[0][0][:36], seg[-1][0][:36]))
        start = idx + 1
