#!/bin/bash
# mixed feature model through bench.py (both arms), and the default headline line once more after the bench.py change
TAG=${1:-r2ad}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --workload synthetic_65536x128+16x50_mixed --steps 10 --warmup 3 > $OUT/${TAG}_bench_mixed.json 2> $OUT/${TAG}_bench_mixed.err; echo "mixed bench rc=$?"; tail -2 $OUT/${TAG}_bench_mixed.err
timeout 600 python bench.py --impl reference --workload synthetic_65536x128+16x50_mixed --steps 3 --warmup 1 > $OUT/${TAG}_bench_mixed_reference_arm.json 2> $OUT/${TAG}_bench_mixed_reference_arm.err; echo "mixed ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_mixed","bench_mixed_reference_arm","bench_n1"):
    try:
        l=json.loads(open("$OUT/${TAG}_%s.json"%f).read().strip().split("\n")[-1])
        print(f, round(l["value"]/1e9,3), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l.get("phase_ms",{}).items()}, (l.get("roofline") or {}).get("frac"), (l.get("production") or {}).get("ms_per_step"), (l.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
