#!/bin/bash
# round 2, final 1-GPU evidence: smoke, the whole GPU suite, BASELINE configs[0]/[1] on both arms, bench lines (headline, reference
# arm, CPHD shard, 1M particles), launch list, DRAM traffic of the update kernel, full captures of update / merge / CPHD update
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu --maxfail=10 --tb=short -q -p no:cacheprovider --timeout 300 > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
timeout 600 python profiles/text_configs.py --impl both > $OUT/${TAG}_text_configs.jsonl 2> $OUT/${TAG}_text_configs.err; cut -c1-230 $OUT/${TAG}_text_configs.jsonl
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().split("\n")[-1])
print({k:l[k] for k in ("value","ms_per_step","phase_ms","gpu_launches")}, l["production"]["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"], l["cpu_baseline"]["value"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cut -c1-200 $OUT/${TAG}_bench_ref.json
timeout 300 python bench.py --workload synthetic_131072x128x50_cphd --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cphd.json 2> $OUT/${TAG}_bench_cphd.err
grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {"update": [0-9.]*, "merge": [0-9.]*' $OUT/${TAG}_bench_cphd.json | head -2
timeout 400 python bench.py --workload synthetic_1048576x128x50_phd --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_1m.json 2> $OUT/${TAG}_bench_1m.err
grep -o '"ms_per_step": [0-9.]*\|"phase_ms": {"update": [0-9.]*, "merge": [0-9.]*' $OUT/${TAG}_bench_1m.json | head -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'update_kernel' --launch-skip 3 --launch-count 1 --csv \
  --log-file $OUT/${TAG}_update_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_traffic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_kernel|merge_fast_kernel' --launch-skip 6 --launch-count 2 \
  -o $OUT/${TAG}_update_merge -f python bench.py --workload synthetic_8192x256x64_phd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_um.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_kernel' --launch-skip 3 --launch-count 1 \
  -o $OUT/${TAG}_cphd_update -f python bench.py --workload synthetic_16384x128x50_cphd --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_cphd.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
