#!/usr/bin/env bash
# ON THE GPU BOX: ncu --set full of kernels matching REGEX in the late-step window of tools/prof_late.py
set -uo pipefail
TAG=${1:-r01}; REGEX=${2:-vein_collisions_listed}; STEPS=${3:-150}; NAME=${4:-late}
O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $STEPS -c ${5:-1} -f -o $O/prof_$NAME \
   python tools/prof_late.py 1000000 $STEPS > $O/ncu_$NAME.log 2>&1
tail -3 $O/ncu_$NAME.log | cut -c1-200
