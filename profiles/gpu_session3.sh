#!/bin/bash
# tests + CPHD shard bench + a full ncu capture of the CPHD update kernel on a 16 384-particle shard
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu --maxfail=12 --tb=short -q > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
timeout 900 python bench.py --workload synthetic_131072x128x50_cphd --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cphd.json 2> $OUT/${TAG}_bench_cphd.err
cat $OUT/${TAG}_bench_cphd.json | cut -c1-1300; tail -3 $OUT/${TAG}_bench_cphd.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'update_kernel' --launch-skip 3 --launch-count 1 \
  -o $OUT/${TAG}_cphd_update -f python bench.py --workload synthetic_16384x128x50_cphd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_cphd.log 2>&1
tail -2 $OUT/${TAG}_ncu_cphd.log
