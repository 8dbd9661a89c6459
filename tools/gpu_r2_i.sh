#!/bin/bash
mkdir -p gpurun_out/r02i
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_rank_slab" > gpurun_out/r02i/test.log 2>&1; rc=$?; echo "slab test exit $rc"; tail -5 gpurun_out/r02i/test.log
[ $rc -ne 0 ] && exit 1
SKIP_WEAK=1 bash tools/gpu_r2_e.sh 2
