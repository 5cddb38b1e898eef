#!/bin/bash
# Second-tier measurements: CPHD shard (BASELINE configs[3] per GPU) and the streaming / resampling shape (configs[4]).
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --workload synthetic_131072x128x50_cphd --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cphd.json 2> $OUT/${TAG}_bench_cphd.err
cat $OUT/${TAG}_bench_cphd.json; tail -3 $OUT/${TAG}_bench_cphd.err
timeout 900 python bench.py --workload synthetic_262144x128x100_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_stream.json 2> $OUT/${TAG}_bench_stream.err
cat $OUT/${TAG}_bench_stream.json; tail -3 $OUT/${TAG}_bench_stream.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_cphd.csv \
  python bench.py --workload synthetic_131072x128x50_cphd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_cphd.log 2>&1
