#!/bin/bash
# quick 1-GPU iteration: merge / update parity tests, then the headline bench line (phases only)
TAG=${1:-it}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_cphd_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 300 -k "not cli" > $OUT/${TAG}_tests.log 2>&1; tail -2 $OUT/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {[^}]*}' $OUT/${TAG}_bench.json | head -3
if [ -n "$2" ]; then timeout 300 python bench.py --workload $2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench2.json 2> $OUT/${TAG}_bench2.err; grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {[^}]*}' $OUT/${TAG}_bench2.json | head -3; fi
