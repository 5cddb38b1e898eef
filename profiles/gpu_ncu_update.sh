#!/bin/bash
# full ncu capture (with source) of the GM-PHD update kernel on an ncu-sized scene
TAG=${1:-rX}
WL=${2:-synthetic_16384x128x50_phd}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'update_kernel' --launch-skip 3 --launch-count 1 \
  -o $OUT/${TAG}_update -f python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_update.log 2>&1
tail -2 $OUT/${TAG}_ncu_update.log
