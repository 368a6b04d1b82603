"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: launch_summary.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] == "ns" else v / 1e3 if r[iu] == "us" else v
    a = agg.setdefault(r[ik].split("(")[0][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-72s %5s %10s %6s" % ("kernel", "n", "total ms", "share"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-72s %5d %10.3f %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
print("%-72s %5d %10.3f" % ("all", sum(a[0] for a in agg.values()), tot))
