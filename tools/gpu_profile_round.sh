#!/usr/bin/env bash
# ON THE GPU BOX: tests + smoke + bench + launch list (gpu_round.sh), then one `ncu --set full` capture per
# contract kernel (springs, particle collisions, vein collisions, finalize) of the same bench command.
set -uo pipefail
TAG=${1:-r01p}
O=gpurun_out/$TAG; mkdir -p $O
bash tools/gpu_round.sh $TAG
for K in ${KERNELS:-springs_kernel particle_collisions_kernel wall_filter_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 1 -f -o $O/full_$K \
     python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full_$K.log 2>&1
  tail -1 $O/ncu_full_$K.log | cut -c1-200
done
ls -la $O
