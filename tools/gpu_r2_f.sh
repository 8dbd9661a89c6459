#!/bin/bash
mkdir -p gpurun_out/r02f
timeout -k 10 420 python -m pytest tests -m gpu -q -x > gpurun_out/r02f/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/r02f/gpu_tests.log
timeout -k 10 200 python bench.py --steps 100 --warmup 10 > gpurun_out/r02f/bench_1gpu.json 2> gpurun_out/r02f/bench_1gpu.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'],d['ms_per_step'],d['e2e']['value'],{k:(round(v['frac'],3),round(v['ms_per_launch']*1e3,1)) for k,v in d['roofline']['contract_kernels'].items()})
PY
