#!/bin/bash
# per-launch durations of the dynamic-map kernels (ncu launch list) at the 65536 x (128 + 16) x 50 mixed shape
TAG=${1:-r2q2}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"dyn_|update_mixed|merge_fast" -c 24 --csv --log-file $OUT/${TAG}_mixed_launches.csv python profiles/mixed_timing.py --particles 65536 --static 128 --dynamic 16 --meas 50 --steps 2 --warmup 1 --oracle-particles 64 > $OUT/${TAG}_mixed_ncu.log 2>&1
echo rc=$?
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/${TAG}_mixed_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((r[ii], r[ki][:40]), {})[r[mi]]=r[vi]
for k,v in list(out.items())[:24]:
    print(k, v)
PY
