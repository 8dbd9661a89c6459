#!/bin/bash
mkdir -p gpurun_out/r02d
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_rank_slab or fused_run or switch" > gpurun_out/r02d/test.log 2>&1
echo "tests exit $?"; tail -15 gpurun_out/r02d/test.log
for rw in "3 8" "0 2"; do SLAB_PROFILE=0 timeout 200 python tools/slab_rank_bench.py $rw 1000000 50 2>&1 | tail -3 | tee -a gpurun_out/r02d/slab_rank.txt; done
timeout 300 python tools/slab_rank_bench.py 3 8 8000000 50 2>&1 | tail -30 | tee -a gpurun_out/r02d/slab_rank_weak.txt
