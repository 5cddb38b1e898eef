#!/bin/bash
# One gpurun session: GPU parity tests, the bench line, the ncu launch list of the same command and the
# full captures the roofline numbers cite.  usage: profiles/gpu_session.sh TAG [skiptests]
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "$2" != "skiptests" ]; then
  timeout 900 python -m pytest tests -m gpu --maxfail=12 --tb=short -q > $OUT/${TAG}_tests.log 2>&1
  tail -5 $OUT/${TAG}_tests.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cat $OUT/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
# full captures on the 8192-particle shape (same C, M): third launch of each kernel = after the merge capacity adapted
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'update_kernel|merge_fast_kernel' --launch-skip 4 --launch-count 2 \
  -o $OUT/${TAG}_update_merge -f python bench.py --workload synthetic_8192x256x64_phd --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
# DRAM traffic of the update kernel on the benchmarked shape
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:update_kernel --launch-skip 3 --launch-count 1 \
  --csv --log-file $OUT/${TAG}_update_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_traffic.log 2>&1
tail -3 $OUT/${TAG}_update_traffic.csv
