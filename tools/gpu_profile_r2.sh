#!/usr/bin/env bash
# ON THE GPU BOX (round 2): the evidence under profiles/r02_*: GPU tests, smoke, bench lines of every BASELINE workload
# that fits one GPU (+ the reference-semantics line), the ncu launch list and one `ncu --set full` capture per contract
# kernel of the same workload.  Every step runs under its own timeout.
set -uo pipefail
TAG=${1:-r02p}
O=gpurun_out/$TAG; mkdir -p $O
make -C oracle >/dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
if [ "${TESTS:-1}" = 1 ]; then
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -12 | tee $O/pytest_gpu.txt
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
fi
echo "== bench long_vein (default)"; timeout 600 python bench.py --steps 100 --warmup 10 2> $O/bench.err | tail -1 > $O/bench_1gpu.json; python tools/show_bench.py $O/bench_1gpu.json | head -26
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>> $O/bench.err | tail -1 > $O/bench_reference_arm.json; cut -c1-300 $O/bench_reference_arm.json
for W in ${WORKLOADS:-cfg1:1000 cfg2:300 cfg3:50}; do
  name=${W%%:*}; steps=${W##*:}
  echo "== bench $name"; timeout 900 python bench.py --workload $name --steps $steps --warmup 10 2>> $O/bench.err | tail -1 > $O/bench_$name.json
  python tools/show_bench.py $O/bench_$name.json | head -8
done
echo "== bench long_vein, reference semantics"; timeout 600 python bench.py --semantics reference --steps 50 --warmup 10 2>> $O/bench.err | tail -1 > $O/bench_refsem.json; python tools/show_bench.py $O/bench_refsem.json | head -8
if [ "${NCU:-1}" = 1 ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-parity > $O/ncu_bench.log 2>&1
  python tools/launch_summary.py $O/launches.csv | tee $O/launches_summary.txt | head -24
  for K in ${KERNELS:-cell_pass_kernel pair_search_kernel row_order_kernel pair_force_kernel}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-6} -c 1 -f -o $O/full_$K \
       python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-parity > $O/ncu_full_$K.log 2>&1
    python tools/ncu_summary.py $O/full_$K.ncu-rep > $O/ncu_full_$K.txt 2>&1
    head -4 $O/ncu_full_$K.txt | cut -c1-160
  done
fi
ls $O | head -40
