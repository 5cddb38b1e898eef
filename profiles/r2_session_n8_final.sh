#!/bin/bash
# round 2, last 8-GPU session with the frozen kernels: weak scaling N = 1, 2, 4, 8 of the headline shape on one box, configs[3]
# (CPHD) and configs[4] (16.7 M particles, strong) at N = 8
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
run() { # N, name, args, port
  if [ "$1" = "1" ]; then
    timeout 600 python bench.py --gpus 1 $3 --no-cpu-baseline > $OUT/${TAG}_$2.json 2> $OUT/${TAG}_$2.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $1 $3 --no-cpu-baseline > $OUT/${TAG}_$2.json 2> $OUT/${TAG}_$2.err
  fi
  python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_$2.json").read().strip().split("\n")[-1])
    print("$2", round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, (l.get("exchange_check") or {}).get("ok"), round(l["production"]["ms_per_step"],3))
except Exception as e:
    print("$2 FAILED", e); print(open("$OUT/${TAG}_$2.err").read()[-1200:])
PY
}
run 1 bench_n1 "--steps 10 --warmup 3" 29540
run 2 bench_n2 "--steps 10 --warmup 3" 29541
run 4 bench_n4 "--steps 10 --warmup 3" 29542
run 8 bench_n8 "--steps 10 --warmup 3" 29543
run 8 bench_n8_cphd "--workload synthetic_131072x128x50_cphd --steps 5 --warmup 3" 29544
run 8 bench_n8_strong16m "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29545
run 4 bench_n4_strong16m "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29546
