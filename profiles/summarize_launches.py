"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:70] if r[4].startswith(("drex", "fp64")) else r[4][:70]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    us = int(r[-1]) / 1e3
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "share_pct", "max_us"])
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, a[0], round(a[1], 1), round(100 * a[1] / tot, 2), round(a[2], 1)])
