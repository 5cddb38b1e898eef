#!/bin/bash
# round 2, last 1-GPU session: the whole GPU suite with the mixed feature model in the library, compute-sanitizer (memcheck and
# racecheck) over the mixed-model tests, the headline bench line and the reference arm once more
TAG=${1:-r2z}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu --maxfail=10 --tb=short -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
SEL='golden_cases or filter_steps or snapshot'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mixed_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT/${TAG}_mixed_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_mixed_memcheck.log; tail -3 $OUT/${TAG}_mixed_memcheck.log | cut -c1-200
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_mixed_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT/${TAG}_mixed_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_mixed_racecheck.log; tail -3 $OUT/${TAG}_mixed_racecheck.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().split("\n")[-1])
print({k:l[k] for k in ("value","ms_per_step","phase_ms","gpu_launches")}, l["roofline"]["frac"], l["e2e"]["value"], l["cpu_baseline"]["value"], l["production"]["ms_per_step"])
PY
