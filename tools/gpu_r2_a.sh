#!/usr/bin/env bash
# ON THE GPU BOX (round 2, call A): the full GPU test suite incl. the new scale tests on the round-1 kernels, the bench
# line with its parity block, and the reference CUDA build on the 100 k long-vein scene.
set -uo pipefail
O=gpurun_out/r2a; mkdir -p $O
make -C oracle >/dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --no-header -rf --durations=8 2>&1 | tail -60 | tee $O/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 2> $O/bench.err | tail -1 > $O/bench.json; cut -c1-1500 $O/bench.json; tail -3 $O/bench.err
echo "== reference CUDA build, long vein 100k"
if [ -x oracle/_ref/ref_headless_longvein100000 ]; then
  timeout 600 oracle/_ref/ref_headless_longvein100000 oracle/_ref/state_longvein100000_seed.bcsd $O bench 100 20 2>&1 | grep -v "out of grid" | tail -2 | tee $O/ref_cuda_longvein100000.txt
fi
