#!/bin/bash
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu --maxfail=12 --tb=short -q > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cut -c1-1100 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
timeout 900 python bench.py --workload synthetic_1048576x128x50_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_1m.json 2> $OUT/${TAG}_bench_1m.err
cut -c1-1100 $OUT/${TAG}_bench_1m.json; tail -3 $OUT/${TAG}_bench_1m.err
