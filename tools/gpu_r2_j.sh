#!/bin/bash
# 8-GPU round-2 measurement: parity (bitwise vs one GPU) + strong, weak and cfg5 bench lines.  Stops at the first failure.
N=${1:-8}
O=gpurun_out/r02j_$N; mkdir -p $O
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 110 29512 tests/mgpu_check.py 1000000 40 > $O/mgpu_1m.log 2>&1; rc=$?; echo "mgpu 1M exit $rc"; grep -E "MGPU_CHECK|step 40|owned-only" $O/mgpu_1m.log | cut -c1-1500
[ $rc -ne 0 ] && { tail -20 $O/mgpu_1m.log; exit 1; }
run 120 29513 bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_strong.json 2> $O/bench_strong.err; rc=$?; echo "bench strong exit $rc"
[ $rc -ne 0 ] && { grep -v "^W1018\|^\*\*\*\|OMP_NUM" $O/bench_strong.err | tail -20; exit 1; }
run 170 29514 bench.py --gpus $N --steps 100 --warmup 10 --scaling weak --no-parity > $O/bench_weak.json 2> $O/bench_weak.err; echo "bench weak exit $?"
run 170 29515 bench.py --gpus $N --steps 50 --warmup 10 --workload cfg5 --no-parity > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 exit $?"
python - <<PY
import json
for f in ("bench_strong","bench_weak","bench_cfg5"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity"))
        print("   ", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
