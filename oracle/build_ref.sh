#!/usr/bin/env bash
# TEST INFRASTRUCTURE: compile the UNMODIFIED reference hot-path sources, where they lie under
# $REF (default /root/reference), into a headless binary + a scene dumper.  Outputs go ONLY to
# oracle/_ref/ (git-ignored, but shipped to the GPU box).  Nothing is copied into the repository.
#
#   oracle/build_ref.sh <cfg-name> [config-overlay-dir] [arch: native|jit]
#
# cfg-name            label for the outputs (ref_headless_<cfg>, ref_scene_dump_<cfg>, scene_<cfg>.bcsd)
# config-overlay-dir  optional directory holding replacement src/config/*.hpp files in the reference's
#                     own header format (the README's workflow: the client app overwrites config/ and the
#                     server is recompiled).  The reference src/ is then copied to a scratch dir under
#                     /tmp, the overlay applied there, and the build runs from the scratch copy.
# arch                native (default): SASS for sm_100 + PTX;  jit: compute_50 PTX only (the reference's
#                     own CMAKE_CUDA_ARCHITECTURES 50, JIT-compiled by the driver on the B200)
#
# Flags follow CMakeLists.txt:20-24,145-146 (-std=c++17 --expt-relaxed-constexpr --maxrregcount=40 -rdc=true).
set -euo pipefail
CFG=${1:?cfg name}
OVERLAY=${2:-}
ARCH=${3:-native}
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT"

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF/src not present - skipping (prebuilt files in oracle/_ref are used as they are)"
  exit 0
fi

SRC=$REF/src
if [ -n "$OVERLAY" ]; then
  SCRATCH=/tmp/bcs_ref_$CFG
  rm -rf "$SCRATCH"; mkdir -p "$SCRATCH"
  cp -r "$REF/src" "$SCRATCH/src"
  chmod -R u+w "$SCRATCH/src"
  cp "$OVERLAY"/*.hpp "$SCRATCH/src/config/"
  SRC=$SCRATCH/src
fi

if [ "$ARCH" = jit ]; then
  GEN="-gencode arch=compute_50,code=compute_50"
  SUFFIX=${CFG}_jit
else
  GEN="-gencode arch=compute_100,code=sm_100 -gencode arch=compute_100,code=compute_100"
  SUFFIX=$CFG
fi

INC="-I$SRC -I$REF/Libraries/include -I$ROOT/include"
# REF_CONSTEXPR_LIMIT=<n>: raise the host compiler's constexpr evaluation limits (large generated veins: the reference's
# vein_factory.hpp walks every vertex / triangle at compile time).  Compiler flags only - the sources stay untouched.
LIM=""; NVLIM=""
if [ -n "${REF_CONSTEXPR_LIMIT:-}" ]; then
  LIM="-fconstexpr-loop-limit=$REF_CONSTEXPR_LIMIT -fconstexpr-ops-limit=${REF_CONSTEXPR_LIMIT}00"
  NVLIM="-Xcompiler -fconstexpr-loop-limit=$REF_CONSTEXPR_LIMIT -Xcompiler -fconstexpr-ops-limit=${REF_CONSTEXPR_LIMIT}00"
fi
NV="nvcc -std=c++17 -O3 $GEN -Wno-deprecated-gpu-targets --expt-relaxed-constexpr --maxrregcount=40 -rdc=true -w $NVLIM $INC"
OBJ=/tmp/bcs_ref_obj_$SUFFIX
rm -rf "$OBJ"; mkdir -p "$OBJ"

g++ -std=c++17 -O1 -w $LIM $INC "$HERE/ref_harness/ref_scene_dump.cpp" -o "$OUT/ref_scene_dump_$CFG" &
pids=()
for f in grids/uniform_grid objects/blood_cells objects/vein_triangles objects/vein_neighbors \
         simulation/vein_collisions simulation/vein_end utilities/cuda_vec3; do
  $NV -c "$SRC/$f.cu" -o "$OBJ/$(basename $f).o" &
  pids+=($!)
done
$NV -c "$HERE/ref_harness/ref_headless.cu" -o "$OBJ/ref_headless.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
wait
$NV "$OBJ"/*.o -o "$OUT/ref_headless_$SUFFIX" -lcurand
"$OUT/ref_scene_dump_$CFG" "$OUT/scene_$CFG.bcsd"
# the integration shim compiled inside the reference's header tree must produce the same user-level scene
g++ -std=c++17 -O1 -w $LIM $INC "$HERE/ref_harness/shim_check.cpp" -o "$OUT/shim_check_$CFG" && "$OUT/shim_check_$CFG" "$OUT/shim_scene_$CFG.bcsd"
rm -rf "$OBJ"
echo "build_ref: built $OUT/ref_headless_$SUFFIX and $OUT/scene_$CFG.bcsd"
