"""Summarise an `ncu --csv` launch list (per-launch metrics) per kernel and for the first bounces.
usage: launch_summary.py <launches.csv> [nFirst]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iK, iV, iM, iI = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
byid = collections.OrderedDict()
for r in rows[1:]:
    try:
        byid.setdefault(r[iI], {"k": r[iK].split("(")[0].replace("void ", "")})[r[iM]] = float(r[iV].replace(",", ""))
    except ValueError:
        pass
agg = collections.OrderedDict()
for d in byid.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("smsp__inst_executed.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("%-34s %5s %10s %7s %12s" % ("kernel", "n", "ms", "share", "warp-inst"))
for k, a in agg.items():
    print("%-34s %5d %10.3f %6.1f%% %12.3e" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, a[2]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("first launches:")
for i, d in list(byid.items())[:n]:
    print("  %3s %-30s %9.1f us  inst %.3e  lanes %.1f  issue %.1f%%" % (
        i, d["k"], d.get("gpu__time_duration.sum", 0) / 1e3, d.get("smsp__inst_executed.sum", 0),
        d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0),
        d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)))
