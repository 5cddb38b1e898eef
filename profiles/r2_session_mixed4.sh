#!/bin/bash
# ncu --set full with source of dyn_update_kernel (mixed feature model), 8192 x (128 + 16) x 50
TAG=${1:-r2v}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"dyn_update" -s 1 -c 1 -o $OUT/${TAG}_dyn_update -f python profiles/mixed_timing.py --particles 8192 --static 128 --dynamic 16 --meas 50 --steps 1 --warmup 1 --oracle-particles 64 > $OUT/${TAG}_ncu.log 2>&1
echo rc=$?; ls -la $OUT/${TAG}_dyn_update.ncu-rep
