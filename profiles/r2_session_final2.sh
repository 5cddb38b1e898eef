#!/bin/bash
# round 2, final evidence for the final kernels: sanitizers, GPU suite, bench lines, captures
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
SEL='not at_scale and not full_size and not cli and not shim and not dist and not accuracy and not step_loop'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_cphd_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL and (full_update or merge_kernels or dense_update or update_modes or resampl or cphd_update)" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log; tail -3 $OUT/${TAG}_racecheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_cphd_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log; tail -3 $OUT/${TAG}_memcheck.log | cut -c1-200
bash profiles/r2_session_final1.sh $TAG
