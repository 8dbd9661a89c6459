#!/usr/bin/env bash
# ON THE GPU BOX (round 2, call B): tests of the new kernels, bench variants.
set -uo pipefail
TAG=${1:-r2b}
O=gpurun_out/$TAG; mkdir -p $O
make -C oracle >/dev/null 2>&1
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q --no-header -rf -x ${PYTEST_ARGS:-} 2>&1 | tail -40 | tee $O/pytest_gpu.txt
for V in ${VARIANTS:-default}; do
  echo "== bench $V"
  case $V in
    default) E="" ;;
    *) E=$(echo $V | tr ',' ' ') ;;
  esac
  env $E timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline ${BENCH_ARGS:-} 2> $O/bench_$V.err | tail -1 > $O/bench_$V.json
  tail -2 $O/bench_$V.err
  python tools/show_bench.py $O/bench_$V.json 2>/dev/null | head -40
done
