#!/usr/bin/env bash
# ON A MULTI-GPU BOX: correctness check + bench at N ranks.  usage: gpu_scale.sh <ngpus> [tag]
set -uo pipefail
N=${1:-2}; TAG=${2:-scale}; O=gpurun_out/$TAG; mkdir -p $O
PORT=$((29500 + N))
if [ "${CHECK:-1}" = 1 ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT tests/mgpu_check.py ${CHECK_PARTICLES:-80000} 40 > $O/mgpu_check_$N.log 2>&1
grep "step\|MGPU\|rror" $O/mgpu_check_$N.log | cut -c1-260 | tail -6
fi
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+20)) bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_g$N.log 2>&1
grep "^{" $O/bench_g$N.log > $O/bench_g$N.json; grep -i "error\|Traceback" $O/bench_g$N.log | head -5
python tools/show_bench.py $O/bench_g$N.json | head -30
if [ "${WEAK:-0}" = 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+40)) bench.py --gpus $N --steps 100 --warmup 10 --scaling weak > $O/bench_weak_g$N.log 2>&1
grep "^{" $O/bench_weak_g$N.log > $O/bench_weak_g$N.json; grep -i "error\|Traceback" $O/bench_weak_g$N.log | head -5
python tools/show_bench.py $O/bench_weak_g$N.json 2>/dev/null | head -30
fi
