#!/usr/bin/env bash
# Runs ON THE GPU BOX (via gpurun): executes the headless reference build (oracle/_ref) on the seeded
# states and brings stage-by-stage dumps back under gpurun_out/ref/.  tools/make_goldens.py then
# turns them into the committed fixtures under tests/golden/.
set -uo pipefail
R=oracle/_ref
O=gpurun_out/ref
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv | tee $O/gpu.txt
run() { # cfg variant mode nsteps dumps...
  local cfg=$1 var=$2 mode=$3 n=$4; shift 4
  local bin=${BIN:-$R/ref_headless_$cfg}
  local d=$O/${TAG:-$cfg}_${var}_${mode}
  mkdir -p $d
  echo "== $bin $cfg $var $mode $n $*"
  timeout 600 $bin $R/state_${cfg}_${var}.bcsd $d $mode $n "$@" > $d/log.txt 2>&1
  echo "exit=$?"; tail -3 $d/log.txt
  grep -c "out of grid" $d/log.txt || true
}
run mini3 spawn staged 100 1 2 3 10 50 100
run mini3 wide staged 100 1 2 3 5 10 20 50 100
run mini3 wide plain 100 1 2 100
run cfg1 spawn staged 100 1 2 10 100
run cfg1 wide staged 100 1 2 50
run cfg1 spawn plain 100 100
BIN=$R/ref_headless_cfg1_jit TAG=cfg1jit run cfg1 spawn staged 100 1 100
run cfg2 wide staged 10 1
# B1 baseline: the reference's own CUDA code on this box (20 warm-up steps + timed steps)
for c in "cfg1 spawn 1000" "cfg2 wide 300" "cfg3 vein 100"; do
  set -- $c
  echo "== bench $1"
  timeout 900 $R/ref_headless_$1 $R/state_$1_$2.bcsd $O bench $3 20 2>&1 | grep -v "out of grid" | tail -2 | tee -a $O/ref_cuda_bench.txt
done
echo "== bench cfg1 jit"; timeout 600 $R/ref_headless_cfg1_jit $R/state_cfg1_spawn.bcsd $O bench 1000 20 2>&1 | tail -1 | tee -a $O/ref_cuda_bench.txt
( cd $O && for d in */; do tar czf ${d%/}.tgz $d && rm -rf $d; done; ls -la )
du -sh gpurun_out
