"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/launch_summary.py gpurun_out/launches.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
iN, iV, iM = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    n = r[iN].split("(")[0][:72]
    a = agg.setdefault(n, [0, 0.0, []])
    a[0] += 1
    a[1] += float(r[iV])
    a[2].append(float(r[iV]))
tot = sum(a[1] for a in agg.values())
print("%-74s %5s %11s %7s %9s %9s" % ("kernel", "n", "total us", "share", "median us", "max us"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    v = sorted(a[2])
    print("%-74s %5d %11.1f %6.1f%% %9.1f %9.1f" % (n, a[0], a[1] / 1e3, 100 * a[1] / tot, v[len(v) // 2] / 1e3, v[-1] / 1e3))
