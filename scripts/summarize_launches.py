"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total)."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    a = agg.setdefault(row["Kernel Name"], [0, 0.0])
    a[0] += 1
    a[1] += v
    n += 1
tot = sum(v[1] for v in agg.values())
print("# %s: %d launches, total %.3f ms (cold-cache, serialised: compare SHARES)" % (path, n, tot / 1e6))
print("%10s %6s %7s  kernel" % ("avg_us", "count", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%10.1f %6d %6.1f%%  %s" % (v[1] / v[0] / 1e3, v[0], 100 * v[1] / tot, k[:110]))
