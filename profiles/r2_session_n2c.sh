#!/bin/bash
# round 2, third 2-GPU session: the whole GPU suite after the deferred state allocation, then BASELINE configs[4] at N = 2
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu --maxfail=10 --tb=short -q -p no:cacheprovider --timeout 300 > $OUT/${TAG}_tests.log 2>&1; tail -3 $OUT/${TAG}_tests.log
run() { # name, env, args, port
  timeout 1200 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 $3 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_$1.json").read().strip().split("\n")[-1])
    print("$1", round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, l.get("exchange_check"), l.get("exchange"), l["production"]["ms_per_step"])
except Exception as e:
    print("$1 FAILED", e); print(open("$OUT/${TAG}_$1.err").read()[-1500:])
PY
}
run bench_n2_strong16m PHDSLAM_MBOX=1 "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29515
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $OUT/${TAG}_mem.txt
