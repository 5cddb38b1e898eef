#!/bin/bash
# CPHD: tests, the configs[3] shard bench, full ncu capture (with source) of the CPHD update kernel
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python bench.py --workload synthetic_16384x128x50_cphd --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_quick.json 2> $OUT/${TAG}_quick.err || { echo "quick bench failed"; tail -3 $OUT/${TAG}_quick.err; exit 1; }
timeout 400 python -m pytest tests -m gpu --maxfail=6 --tb=short -q -p no:cacheprovider --timeout 120 > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
timeout 300 python bench.py --workload synthetic_131072x128x50_cphd --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cphd.json 2> $OUT/${TAG}_bench_cphd.err
cut -c1-1300 $OUT/${TAG}_bench_cphd.json; tail -3 $OUT/${TAG}_bench_cphd.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_kernel' --launch-skip 3 --launch-count 1 \
  -o $OUT/${TAG}_cphd_update -f python bench.py --workload synthetic_16384x128x50_cphd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_cphd.log 2>&1
tail -2 $OUT/${TAG}_ncu_cphd.log
