#!/bin/bash
# round 2, 1-GPU session B: tests after the merge / fused-mode / shim changes, racecheck, bench lines (headline, CPHD shard,
# 1M particles), launch list, DRAM traffic of the update kernel (headline and 1M workloads), full captures of update / merge
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu --maxfail=10 --tb=short -q -p no:cacheprovider --timeout 300 > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
SEL='not at_scale and not full_size and not cli and not shim and not dist and not accuracy and not step_loop'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_cphd_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL and (full_update or merge_kernels or dense_update or update_modes or resampl or cphd_update)" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log; tail -3 $OUT/${TAG}_racecheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_cphd_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log; tail -3 $OUT/${TAG}_memcheck.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().split("\n")[-1])
print({k:l[k] for k in ("value","ms_per_step","phase_ms","production","gpu_launches")}, l["roofline"]["frac"], l["e2e"])
PY
timeout 300 python bench.py --workload synthetic_131072x128x50_cphd --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cphd.json 2> $OUT/${TAG}_bench_cphd.err
grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {"update": [0-9.]*, "merge": [0-9.]*\|"production": {[^}]*}' $OUT/${TAG}_bench_cphd.json | head -3
timeout 400 python bench.py --workload synthetic_1048576x128x50_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_1m.json 2> $OUT/${TAG}_bench_1m.err
grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {"update": [0-9.]*, "merge": [0-9.]*\|"production": {[^}]*}' $OUT/${TAG}_bench_1m.json | head -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'update_kernel' --launch-skip 3 --launch-count 1 --csv \
  --log-file $OUT/${TAG}_update_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_traffic.log 2>&1
tail -3 $OUT/${TAG}_update_traffic.csv | cut -d, -f13-15
# the 1M-particle workload streams through the update buffer in 7 batches: one launch per batch, all captured
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'update_kernel' --launch-skip 7 --launch-count 7 --csv \
  --log-file $OUT/${TAG}_update_traffic_1m.csv python bench.py --workload synthetic_1048576x128x50_phd --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_traffic_1m.log 2>&1
tail -8 $OUT/${TAG}_update_traffic_1m.csv | cut -d, -f13-15
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_kernel|merge_fast_kernel' --launch-skip 6 --launch-count 2 \
  -o $OUT/${TAG}_update_merge -f python bench.py --workload synthetic_8192x256x64_phd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_um.log 2>&1
ls -la $OUT/${TAG}_update_merge.ncu-rep
