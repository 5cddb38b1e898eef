#!/bin/bash
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu --maxfail=12 --tb=short -q > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cat $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
PHDSLAM_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1_serial.json 2> $OUT/${TAG}_bench_n1_serial.err
cut -c1-900 $OUT/${TAG}_bench_n1_serial.json
timeout 900 python bench.py --workload synthetic_262144x128x100_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_stream.json 2> $OUT/${TAG}_bench_stream.err
cut -c1-1200 $OUT/${TAG}_bench_stream.json; tail -3 $OUT/${TAG}_bench_stream.err
