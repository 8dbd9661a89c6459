#!/usr/bin/env bash
# ON THE GPU BOX: `ncu --set full` captures of the named kernels (one report each) during a short bench run + summaries.
set -uo pipefail
TAG=${1:-r2n}; shift
O=gpurun_out/$TAG; mkdir -p $O
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-4} -c 1 -f -o $O/full_$K \
     python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity ${BENCH_ARGS:-} > $O/ncu_$K.log 2>&1
  tail -1 $O/ncu_$K.log | cut -c1-160
  python tools/ncu_summary.py $O/full_$K.ncu-rep > $O/ncu_full_$K.txt 2>&1
  head -22 $O/ncu_full_$K.txt
done
