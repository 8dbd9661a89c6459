#!/bin/bash
mkdir -p gpurun_out/r02h
timeout -k 10 420 python -m pytest tests -m gpu -q -x > gpurun_out/r02h/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/r02h/gpu_tests.log
timeout -k 10 150 python bench.py --workload cfg2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02h/cfg2.json 2> gpurun_out/r02h/cfg2.err; echo "cfg2 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h/cfg2.json').read().strip().splitlines()[-1])
print(d['value'],d['ms_per_step'],d['e2e']['value'],list(d['kernels'].keys())[:6], d['parity']['ok'])
PY
