#!/bin/bash
# usage: gpu_r2_e.sh <ranks> [weak]   multi-GPU parity (bitwise vs one GPU) + strong (/ weak) bench lines.  Stops at the first failure.
N=${1:-2}
O=gpurun_out/r02e_$N; mkdir -p $O
run() { timeout -k 5 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
if [ -z "$SKIP_MGPU" ]; then
run 70 29511 tests/mgpu_check.py 40000 60 > $O/mgpu_40k.log 2>&1; rc=$?; echo "mgpu 40k exit $rc"; grep -E "MGPU_CHECK|step 60" $O/mgpu_40k.log | cut -c1-700
[ $rc -ne 0 ] && { tail -20 $O/mgpu_40k.log; exit 1; }
run 100 29512 tests/mgpu_check.py 1000000 40 > $O/mgpu_1m.log 2>&1; rc=$?; echo "mgpu 1M exit $rc"; grep -E "MGPU_CHECK|step 40" $O/mgpu_1m.log | cut -c1-900
[ $rc -ne 0 ] && { tail -20 $O/mgpu_1m.log; exit 1; }
fi
run 120 29513 bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_strong.json 2> $O/bench_strong.err; rc=$?; echo "bench strong exit $rc"
[ $rc -ne 0 ] && { tail -20 $O/bench_strong.err; exit 1; }
if [ -n "$2" ]; then
  run 160 29514 bench.py --gpus $N --steps 100 --warmup 10 --scaling weak --no-parity > $O/bench_weak.json 2> $O/bench_weak.err; echo "bench weak exit $?"
fi
python - <<PY
import json
for f in ("bench_strong","bench_weak"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity"))
        print("   ", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
