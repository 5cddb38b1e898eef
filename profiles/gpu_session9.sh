#!/bin/bash
# evidence refresh with the final round-1 kernels: launch list of the bench command, DRAM traffic of the update kernel at
# the headline size, full captures (with source) of update (8192-particle shape) and merge
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
tail -1 $OUT/${TAG}_launches.log | cut -c1-200
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'update_kernel' --launch-skip 3 --launch-count 1 --csv \
  --log-file $OUT/${TAG}_update_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_traffic.log 2>&1
tail -3 $OUT/${TAG}_update_traffic.csv | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_kernel|merge_fast_kernel' --launch-skip 6 --launch-count 2 \
  -o $OUT/${TAG}_update_merge -f python bench.py --workload synthetic_8192x256x64_phd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_um.log 2>&1
tail -2 $OUT/${TAG}_ncu_um.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cut -c1-300 $OUT/${TAG}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cut -c1-300 $OUT/${TAG}_bench_ref.json
