#!/bin/bash
# round 2: mixed feature model -- all its GPU tests (parity, golden, tracking, CLI), then its timing at two shapes
TAG=${1:-r2n}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mixed_gpu.py tests/test_mixed_tracking.py -x -q > $OUT/${TAG}_mixed_tests.log 2>&1; echo "mixed tests rc=$?"; tail -12 $OUT/${TAG}_mixed_tests.log
timeout 600 python profiles/mixed_timing.py --particles 65536 --static 128 --dynamic 16 --meas 50 > $OUT/${TAG}_mixed_timing_65536.json 2> $OUT/${TAG}_mixed_timing_65536.err; echo "timing rc=$?"; tail -c 1200 $OUT/${TAG}_mixed_timing_65536.json; tail -3 $OUT/${TAG}_mixed_timing_65536.err
timeout 600 python profiles/mixed_timing.py --particles 16384 --static 64 --dynamic 48 --meas 32 > $OUT/${TAG}_mixed_timing_16384.json 2> $OUT/${TAG}_mixed_timing_16384.err; echo "timing rc=$?"; tail -c 1200 $OUT/${TAG}_mixed_timing_16384.json; tail -3 $OUT/${TAG}_mixed_timing_16384.err
