#!/bin/bash
# round 2, 1-GPU session A: the whole GPU test suite (new: CPHD vs the reference's kernels, shim replay, update modes,
# configs[1] at 4096 particles), BASELINE configs[0]/[1] on both arms, sanitizer runs, bench lines with the new keys.
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu --maxfail=10 --tb=short -q -p no:cacheprovider --timeout 300 > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
timeout 600 python profiles/text_configs.py --impl both > $OUT/${TAG}_text_configs.jsonl 2> $OUT/${TAG}_text_configs.err; cut -c1-260 $OUT/${TAG}_text_configs.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cut -c1-300 $OUT/${TAG}_bench_n1.json; tail -2 $OUT/${TAG}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cut -c1-200 $OUT/${TAG}_bench_ref.json; tail -2 $OUT/${TAG}_bench_ref.err
# compute-sanitizer on small configurations: memcheck and racecheck over the parity tests
SEL='not at_scale and not full_size and not cli and not shim and not dist and not accuracy and not step_loop'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_cphd_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log; tail -4 $OUT/${TAG}_memcheck.log | cut -c1-200
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL and (full_update or merge_kernels or dense_update or update_modes or resampl)" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log; tail -4 $OUT/${TAG}_racecheck.log | cut -c1-200
