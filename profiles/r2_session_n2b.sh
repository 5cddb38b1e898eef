#!/bin/bash
# round 2, second 2-GPU session: the strong-scaling path (import_tiled, tiled snapshot) at 2 M and at BASELINE configs[4]'s
# 16.7 M particles on 2 GPUs; merge_fast_kernel at 8 CTAs per SM (capacity pinned to 672 candidates)
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
run() { # name, env, args, port
  timeout 900 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 $3 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_$1.json").read().strip().split("\n")[-1])
    print("$1", round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, l.get("exchange_check"), l.get("exchange"), l["production"]["ms_per_step"])
except Exception as e:
    print("$1 FAILED", e); print(open("$OUT/${TAG}_$1.err").read()[-1500:])
PY
}
run bench_n2_strong2m PHDSLAM_MBOX=1 "--workload synthetic_2097152x128x100_phd --steps 3 --warmup 2" 29514
nvidia-smi --query-gpu=memory.used --format=csv > $OUT/${TAG}_mem0.txt
run bench_n2_strong16m PHDSLAM_MBOX=1 "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29515
for cap in; do
  PHDSLAM_MERGE_CAP=$cap timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_cap$cap.json 2> $OUT/${TAG}_bench_cap$cap.err
  grep -o '"phase_ms": {"update": [0-9.]*, "merge": [0-9.]*' $OUT/${TAG}_bench_cap$cap.json
done
