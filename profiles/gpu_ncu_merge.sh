#!/bin/bash
# full ncu capture (with source) of merge_fast_kernel on an 8192-particle scene of the headline shape
TAG=${1:-rX}
WL=${2:-synthetic_8192x256x64_phd}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_fast_kernel' --launch-skip 3 --launch-count 1 \
  -o $OUT/${TAG}_merge_fast -f python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_merge.log 2>&1
tail -2 $OUT/${TAG}_ncu_merge.log
ls -la $OUT/${TAG}_merge_fast.ncu-rep
