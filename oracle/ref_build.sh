#!/bin/bash
# oracle/ref_build.sh -- compile the REFERENCE's own kernels for the CPU (oracle/_ref/libphd_ref.so).
#
# TEST INFRASTRUCTURE.  Runs only where /root/reference exists (this container); the built .so travels to the
# GPU box, /root/reference does not.  No reference source is copied into the repository: the recipe reads the
# sources where they lie and writes only below oracle/_ref/ (git-ignored).
#
# Why not the reference's own build: HEAD does not compile (MotionModel undeclared, thrust headers,
# cuPrintf needs sm_11 headers, `using namespace thrust` vs CCCL) and needs Boost/Eigen/matio/qmake, none of
# which is installed.  What IS compilable is the arithmetic: the __global__ kernels and __device__ helpers of
# the hot path are plain C++ once threadIdx/__syncthreads/__shared__ exist.  oracle/ref_shim/cuda_emul.h
# provides those (one ucontext fiber per CUDA thread), and this script
#   1. takes the hot-path functions out of src/phdfilter.cu and src/main.cpp BY LINE RANGE, verbatim, into
#      oracle/_ref/gen/*.inc (each range is checked against the function name it must start with);
#   2. makes a copy of src/device_math.cuh in which every step of the three warp-synchronous shared-memory
#      reductions (sumByReduction / productByReduction / maxByReduction, :452-531) is followed by __syncwarp()
#      -- the fix every post-Volta port of that code needs; on the hardware the reference was written for the
#      warp ran in lock-step and the result is the same;
#   3. takes the EAP map reduction out of src/gm_reduce.cpp (GaussianX, its Mahalanobis distance, the
#      reduceGaussianMixture template) and compiles it over oracle/ref_shim/eigen3/Eigen/* -- a ~100-line stand-in for
#      the dozen Eigen operations that file uses (Eigen is not installed);
#   4. compiles oracle/ref_harness.cpp (ours: marshalling + the host glue between kernels) with them.
# -ffp-contract=off: the CPU build must not fuse a*b+c on its own; -fpermissive -w: 2012-era C++.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PHD_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
[ -d "$REF/src" ] || { echo "ref_build: $REF/src not found, nothing to do"; exit 0; }
mkdir -p "$GEN"

# extract FILE FIRST LAST NAME  -> appends lines FIRST..LAST of FILE; NAME must occur in the first 3 lines
extract() {
  local file="$1" first="$2" last="$3" name="$4"
  sed -n "${first},$((first + 2))p" "$file" | grep -q "$name" || { echo "ref_build: $file:$first is not $name" >&2; exit 1; }
  [ "$(sed -n "${last}p" "$file")" = "}" ] || { echo "ref_build: $file:$last does not close $name" >&2; exit 1; }
  echo "/* ---- $(basename "$file"):$first-$last ($name) ---- */"
  sed -n "${first},${last}p" "$file"
}

K="$REF/src/phdfilter.cu"
{
  extract "$K" 205 242 computeBirth
  extract "$K" 785 825 phdPredictKernelAckerman
  extract "$K" 827 859 phdPredictKernel
  extract "$K" 867 888 cardinalityPredictKernel
  extract "$K" 1279 1358 computeInRangeKernel
  extract "$K" 1824 1925 preUpdateSynthKernel
  # phdUpdateKernel reads sdata[0] after each per-measurement reduction (`sum += sdata[0]`, :2210) with no
  # barrier before thread 0 overwrites sdata[0] in the next reduction.  On hardware the other warps read it
  # immediately and the race is benign; the emulator runs thread 0 far ahead, so the barrier the kernel's own
  # first loop has (:2184-2185) is added there too.
  extract "$K" 2083 2321 phdUpdateKernel | sed -E 's/^( *sum \+= sdata\[0\] ;)\s*$/\1 __syncthreads() ;/'
  extract "$K" 2707 2898 phdUpdateMergeKernel
} > "$GEN/ref_kernels.inc"
[ "$(grep -c 'sum += sdata\[0\] ; __syncthreads() ;' "$GEN/ref_kernels.inc")" -eq 1 ] || { echo "ref_build: barrier patch failed" >&2; exit 1; }

# ---- mixed feature model (static + constant-velocity features): the 4-D birth and pre-update device functions, the map
# prediction kernel and phdUpdateKernelMixed.  Like phdUpdateKernel above, the mixed kernel reads sdata[0] after the
# per-measurement reduction (`normalizer += sdata[0]`, :2472) with no barrier before thread 0 starts the next reduction:
# the same barrier is added for the emulator.
{
  extract "$K" 244 299 computeBirth
  extract "$K" 301 395 computePreUpdate
  extract "$K" 397 521 computePreUpdate
  extract "$K" 910 963 predictMapKernelMixed
  extract "$K" 2323 2635 phdUpdateKernelMixed | sed -E 's/^( *normalizer \+= sdata\[0\] ;)\s*$/\1 __syncthreads() ;/'
} > "$GEN/ref_mixed.inc"
[ "$(grep -c 'normalizer += sdata\[0\] ; __syncthreads() ;' "$GEN/ref_mixed.inc")" -eq 1 ] || { echo "ref_build: mixed barrier patch failed" >&2; exit 1; }

# ---- CPHD: the reference's multi-object update kernels.  HEAD keeps them COMMENTED OUT (every line prefixed with `//`,
# src/phdfilter.cu:701-748,1430-1822); the live older forms are in src/phdfilter.cu.bak.  Both are made executable here:
#   head: the commented ranges with the ONE leading `//` of every line removed (nested `////` comments stay comments) --
#         the newest CPHD arithmetic the reference holds (linear-domain signed ESF recursion, log|.| at the end);
#   bak : the live kernels of the .bak, verbatim (log-domain ESF recursion whose leave-one-out variant takes fabs() of a
#         DIFFERENCE of exponentials at every step, and <Psi1d,p> normalised by the wrong maximum: see tests/test_ref_pin.py).
# extract_commented FILE FIRST LAST NAME: as extract, for a `//`-commented range
extract_commented() {
  local file="$1" first="$2" last="$3" name="$4"
  sed -n "${first},$((first + 2))p" "$file" | grep -q "^//.*$name" || { echo "ref_build: $file:$first is not commented $name" >&2; exit 1; }
  [ "$(sed -n "${last}p" "$file")" = "//}" ] || { echo "ref_build: $file:$last does not close $name" >&2; exit 1; }
  [ "$(sed -n "${first},${last}p" "$file" | grep -vc '^//')" -eq "$(sed -n "${first},${last}p" "$file" | grep -c '^$')" ] || { echo "ref_build: $file:$first-$last has live lines" >&2; exit 1; }
  echo "/* ---- $(basename "$file"):$first-$last ($name, uncommented) ---- */"
  sed -n "${first},${last}p" "$file" | sed 's#^//##'
}
{
  extract_commented "$K" 543 607 computePreUpdateComponents
  extract_commented "$K" 702 748 cphdConstantsKernel
  extract_commented "$K" 1430 1511 cphdPreUpdateKernel
  extract_commented "$K" 1524 1618 computeEsfKernel
  extract_commented "$K" 1626 1769 computePsiKernel
  extract_commented "$K" 1780 1822 cphdUpdateKernel
} > "$GEN/ref_cphd_head.inc"
B="$REF/src/phdfilter.cu.bak"
{
  extract "$B" 369 415 cphdConstantsKernel
  extract "$B" 1058 1181 cphdPreUpdateKernel
  extract "$B" 1194 1278 computeEsfKernel
  extract "$B" 1286 1428 computePsiKernel
  extract "$B" 1436 1478 cphdUpdateKernel
} > "$GEN/ref_cphd_bak.inc"

# the birth-term host loop is inline in phdUpdateSynth (:3468-3510): taken as a block, wrapped by the harness
sed -n "3468p" "$K" | grep -q "for ( int i = 0 ; i < n_particles ; i++){" || { echo "ref_build: births loop moved" >&2; exit 1; }
sed -n "3470,3475p" "$K" | grep -q "invert measurement" || { echo "ref_build: births loop moved" >&2; exit 1; }
sed -n "3468,3510p" "$K" > "$GEN/ref_births.inc"

M="$REF/src/main.cpp"
{
  extract "$M" 147 167 loadTimestamps
  extract "$M" 169 190 loadControls
  extract "$M" 192 208 parseMeasurements
  extract "$M" 221 245 loadMeasurements
  extract "$M" 247 264 loadTrajectory
  extract "$M" 290 316 computeExpectedMap
  extract "$M" 318 388 recoverSlamState
  extract "$M" 452 501 resampleParticles
  extract "$M" 848 954 writeLog
} > "$GEN/ref_host.inc"

# the EAP map reduction (src/gm_reduce.cpp): GaussianX, mahalanobisDistance(GaussianX, GaussianX), compare_gaussians and
# the reduceGaussianMixture template, verbatim; the Gaussian4D overload of mahalanobisDistance (:39-53, Eigen::Map on
# fixed-size types, never called by the template) is left out.  Eigen is absent: oracle/ref_shim/eigen3/Eigen/* stands in.
G="$REF/src/gm_reduce.cpp"
sed -n "8p" "$G" | grep -q "using namespace Eigen" || { echo "ref_build: gm_reduce.cpp moved" >&2; exit 1; }
sed -n "10p" "$G" | grep -q "struct GaussianX" || { echo "ref_build: gm_reduce.cpp moved" >&2; exit 1; }
sed -n "55p" "$G" | grep -q "compare_gaussians" || { echo "ref_build: gm_reduce.cpp moved" >&2; exit 1; }
{
  echo "/* ---- gm_reduce.cpp:8,10-28 (GaussianX) ---- */"
  sed -n "8p;10,28p" "$G"
  extract "$G" 30 37 mahalanobisDistance
  echo "/* ---- gm_reduce.cpp:55 ---- */"
  sed -n "55p" "$G"
  extract "$G" 57 134 reduceGaussianMixture
} > "$GEN/ref_gm_reduce.inc"

# the input schedule of time-stamped runs is inline in run_synth's loop (src/main.cpp:1188-1230): taken as a block
sed -n "1187p" "$M" | grep -q "if ( has_timestamps )" || { echo "ref_build: event schedule moved" >&2; exit 1; }
sed -n "1231p" "$M" | grep -q "else" || { echo "ref_build: event schedule moved" >&2; exit 1; }
sed -n "1188,1230p" "$M" > "$GEN/ref_events.inc"

# device_math.cuh with __syncwarp() after every warp-synchronous reduction step
sed -E 's/^( *sdata\[tid\] = [a-z_A-Z]+ = .*sdata\[tid *\+ *(32|16|8|4|2|1)\].*;)\s*$/\1 __syncwarp();/' \
  "$REF/src/device_math.cuh" > "$GEN/device_math_syncwarp.cuh"
n=$(grep -c "__syncwarp" "$GEN/device_math_syncwarp.cuh")
[ "$n" -eq 18 ] || { echo "ref_build: expected 18 warp-synchronous steps in device_math.cuh, patched $n" >&2; exit 1; }

g++ -O1 -std=c++14 -fPIC -shared -fpermissive -w -ffp-contract=off \
  -I "$HERE/ref_shim" -I "$GEN" -I "$REF/src" -I "$HERE/../include" \
  -o "$OUT/libphd_ref.so" "$HERE/ref_harness.cpp"
echo "ref_build: built $OUT/libphd_ref.so from $REF"

# ---- the drop-in, executed: the reference's own run_synth (src/main.cpp:1076-1313, verbatim: input loading, particle
# initialisation, the time-step loop with its host resampleParticles) + its recoverSlamState / writeLog / loaders, linked
# with cuda-phdslam_b200/shim/phdfilter_b200.cpp and libphdslam.so (oracle/shim_replay.cpp says what is stubbed).
sed -n "1075p" "$M" | grep -q "void run_synth(bool profile_run){" || { echo "ref_build: run_synth moved" >&2; exit 1; }
sed -n "1178p" "$M" | grep -q "for (int n = 0 ; n < nSteps ; n++ )" || { echo "ref_build: run_synth loop moved" >&2; exit 1; }
sed -n "1314p" "$M" | grep -q "else" || { echo "ref_build: run_synth loop end moved" >&2; exit 1; }
sed -n "1076,1313p" "$M" > "$GEN/ref_run_synth_body.inc"
LIBDIR="$HERE/../cuda-phdslam_b200"
if [ -f "$LIBDIR/libphdslam.so" ]; then
  g++ -O1 -std=c++14 -fpermissive -w -ffp-contract=off \
    -I "$HERE/ref_shim" -I "$GEN" -I "$REF/src" -I "$HERE/../include" \
    -o "$OUT/shim_replay" "$HERE/shim_replay.cpp" "$LIBDIR/shim/phdfilter_b200.cpp" \
    -L "$LIBDIR" -lphdslam -Wl,-rpath,'$ORIGIN/../../cuda-phdslam_b200'
  echo "ref_build: built $OUT/shim_replay (the reference's run_synth over the drop-in shim)"
else
  echo "ref_build: libphdslam.so not built yet, skipping shim_replay"
fi
