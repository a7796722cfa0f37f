#!/usr/bin/env python
"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid.
usage: tools/launch_summary.py <launches.csv> [header line ...]"""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    key = (name, r["Grid Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) / 1e6        # ns -> ms
total = sum(a[1] for a in agg.values())
for h in sys.argv[2:]:
    print(h)
print(f"{len(rows)} launches, {total:.3f} ms of kernel time (serialised, cold cache: compare SHARES)\n")
print(f"{'kernel':<72} {'grid':>16} {'launches':>8} {'total ms':>10} {'avg us':>10} {'share':>7}")
for (name, grid), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:72]:<72} {grid:>16} {n:>8} {ms:>10.3f} {ms / n * 1e3:>10.1f} {100 * ms / total:>6.1f}%")
