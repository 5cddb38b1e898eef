#!/bin/bash
# guarded session: a short sanitizer-free smoke first (a hung kernel costs minutes, not the whole call)
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python bench.py --workload synthetic_1024x64x32_phd --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_quick.json 2> $OUT/${TAG}_quick.err || { echo "quick bench failed"; tail -3 $OUT/${TAG}_quick.err; exit 1; }
timeout 400 python -m pytest tests -m gpu --maxfail=6 --tb=short -q -p no:cacheprovider --timeout 120 > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cut -c1-1100 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
timeout 400 python bench.py --workload synthetic_1048576x128x50_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_1m.json 2> $OUT/${TAG}_bench_1m.err
cut -c1-1100 $OUT/${TAG}_bench_1m.json; tail -3 $OUT/${TAG}_bench_1m.err
