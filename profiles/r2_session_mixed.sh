#!/bin/bash
# round 2: the mixed feature model (feature_model = 2) on the GPU -- its parity tests, then a sanity bench of the headline
# (static) workload, whose kernels must be what they were
TAG=${1:-r2m}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_mixed_gpu.py -x -q > $OUT/${TAG}_mixed_tests.log 2>&1; echo "mixed tests rc=$?"; tail -15 $OUT/${TAG}_mixed_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().split("\n")[-1])
print(round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, l["roofline"]["frac"])
PY
