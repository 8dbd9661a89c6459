#!/bin/bash
mkdir -p gpurun_out/r02c
for rw in "3 8" "0 8" "0 2" "1 4"; do timeout 200 python tools/slab_rank_bench.py $rw 1000000 50 >> gpurun_out/r02c/slab_rank.txt 2>&1; done
timeout 300 python tools/slab_rank_bench.py 3 8 8000000 50 >> gpurun_out/r02c/slab_rank_weak.txt 2>&1
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r02c/bench_1gpu.json 2> gpurun_out/r02c/bench_1gpu.err
cat gpurun_out/r02c/slab_rank.txt gpurun_out/r02c/slab_rank_weak.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c/bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'],d['ms_per_step'],{k:(v['frac'],v['ms_per_launch']) for k,v in d['roofline']['contract_kernels'].items()})
PY
