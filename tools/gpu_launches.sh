#!/usr/bin/env bash
# ON THE GPU BOX: ncu launch list (per-launch durations, cold caches, serialised) of a short bench run + per-kernel summary.
set -uo pipefail
TAG=${1:-r2l}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-60} -c ${COUNT:-200} --csv --log-file $O/launches.csv \
    python tools/quick_bench.py --steps 12 --warmup 3 > $O/ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$O/launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if hdr:
    h = rows[hdr[0]]; data = rows[hdr[0] + 1:]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open("$O/launches_summary.txt", "w") as f:
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            line = f"{k:60s} launches={v[0]:4d} avg_us={v[1]/v[0]/1e3:8.1f} share={v[1]/tot:6.1%}"
            print(line); f.write(line + "\n")
PY
