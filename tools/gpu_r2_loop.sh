#!/bin/bash
# Repeats the GPU parity + scale tests to shake out intermittent failures; then the full GPU suite and smoke().
mkdir -p gpurun_out/loop
for i in 1 2 3 4; do
  timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/loop/run$i.log 2>&1
  echo "run $i exit $?" | tee -a gpurun_out/loop/summary.txt
  tail -2 gpurun_out/loop/run$i.log | tee -a gpurun_out/loop/summary.txt
done
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/loop/full.log 2>&1
echo "full exit $?" | tee -a gpurun_out/loop/summary.txt
tail -3 gpurun_out/loop/full.log | tee -a gpurun_out/loop/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/loop/summary.txt
