#!/usr/bin/env bash
# ON THE GPU BOX: ncu --set full of one kernel (regex) while tools/bench_stage.py <stage> runs; summary to gpurun_out/<tag>/
set -uo pipefail
TAG=$1; STAGE=$2; K=$3
O=gpurun_out/$TAG; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KR:-$K}" -s ${SKIP:-3} -c 1 -f -o $O/full_$K python tools/bench_stage.py $STAGE 6 > $O/ncu_$K.log 2>&1
tail -1 $O/ncu_$K.log | cut -c1-160
python tools/ncu_summary.py $O/full_$K.ncu-rep > $O/ncu_full_$K.txt 2>&1
python tools/ncu_lines.py $O/full_$K.ncu-rep 30 > $O/ncu_lines_$K.txt 2>&1
python tools/ncu_lines.py $O/full_$K.ncu-rep 30 --by-instructions > $O/ncu_ins_$K.txt 2>&1
head -24 $O/ncu_full_$K.txt
