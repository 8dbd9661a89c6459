#!/usr/bin/env bash
# ON THE GPU BOX: one `ncu --set full` capture of the named kernels (regex) during a short bench run.
set -uo pipefail
TAG=${1:-r01}; REGEX=${2:-vein_collisions_kernel}; SKIP=${3:-6}; COUNT=${4:-1}
O=gpurun_out/$TAG; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -f -o $O/prof_${5:-k} \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log | cut -c1-200
ls -la $O
