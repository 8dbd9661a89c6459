#!/bin/bash
mkdir -p gpurun_out/r02g
for g in rows cells; do
  BCS_GRID=$g timeout -k 10 150 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02g/cfg2_$g.json 2> gpurun_out/r02g/cfg2_$g.err; echo "cfg2 $g exit $?"
done
timeout -k 10 300 python bench.py --workload cfg5 --steps 30 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r02g/cfg5_1gpu.json 2> gpurun_out/r02g/cfg5_1gpu.err; echo "cfg5 exit $?"
python - <<'PY'
import json
for f in ("cfg2_rows","cfg2_cells","cfg5_1gpu"):
    try:
        d=json.loads(open(f"gpurun_out/r02g/{f}.json").read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/r02g/cfg5_1gpu.err
