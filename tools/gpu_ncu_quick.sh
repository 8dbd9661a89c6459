#!/usr/bin/env bash
# ON THE GPU BOX: ncu --set full of one kernel (regex) while tools/quick_bench.py runs; summaries to gpurun_out/<tag>/
set -uo pipefail
TAG=$1; K=$2
O=gpurun_out/$TAG; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-5} -c 1 -f -o $O/full_$K python tools/quick_bench.py --steps 12 --warmup 3 > $O/ncu_$K.log 2>&1
tail -1 $O/ncu_$K.log | cut -c1-160
python tools/ncu_summary.py $O/full_$K.ncu-rep > $O/ncu_full_$K.txt 2>&1
python tools/ncu_lines.py $O/full_$K.ncu-rep 30 --by-instructions > $O/ncu_ins_$K.txt 2>&1
head -36 $O/ncu_full_$K.txt | cut -c1-230
