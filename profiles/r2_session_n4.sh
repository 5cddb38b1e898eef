#!/bin/bash
# round 2, 4-GPU session: headline shape (weak) and BASELINE configs[4] (strong, 16.7 M particles) at N = 4
TAG=${1:-r2i}
N=4
OUT=gpurun_out
mkdir -p $OUT
run() { # name, env, args, port
  timeout 900 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $4 bench.py --gpus $N $3 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_$1.json").read().strip().split("\n")[-1])
    ex=l.get("exchange") or {}
    print("$1", round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, (l.get("exchange_check") or {}).get("ok"), (l.get("exchange_check") or {}).get("migrated_checked"), round(ex.get("achieved_GBps_lower_bound") or 0,1), "GB/s", round(l["production"]["ms_per_step"],3))
except Exception as e:
    print("$1 FAILED", e); print(open("$OUT/${TAG}_$1.err").read()[-1200:])
PY
}
run bench_n4 PHDSLAM_MBOX=1 "--steps 10 --warmup 3" 29531
run bench_n4_strong16m PHDSLAM_MBOX=1 "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29532
