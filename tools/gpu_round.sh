#!/usr/bin/env bash
# Runs ON THE GPU BOX: parity tests, smoke, bench and the ncu launch list.  Outputs under gpurun_out/.
set -uo pipefail
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
make -C oracle >/dev/null 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-100} --warmup 10 2>&1 | tail -2 | tee $O/bench.json
if [ "${NCU:-1}" = 1 ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv --log-file $O/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
  tail -2 $O/ncu_bench.log | cut -c1-300
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
            line = f"{k:60s} launches={v[0]:4d} total_us={v[1]/1e3:10.1f} share={v[1]/tot:6.1%}"
            print(line); f.write(line + "\n")
PY
fi
