#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if not hdr:
    sys.exit("no launch table in " + sys.argv[1])
h = rows[hdr[0]]; data = rows[hdr[0] + 1:]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} launches={v[0]:4d} avg_us={v[1] / v[0] / 1e3:8.1f} share={v[1] / tot:6.1%}")
