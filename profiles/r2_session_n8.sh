#!/bin/bash
# round 2, N-GPU session (N = 8 or 4): sharding parity tests incl. the skewed 3 / 4 / 8-GPU cases, then bench lines with
# exchange_check: headline shape (weak), CPHD configs[3] (N = 8: its literal size), configs[4] strong scaling at 16.7 M particles
N=${1:-8}
TAG=${2:-r2h}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -v --timeout 300 > $OUT/${TAG}_dist_tests.log 2>&1
echo "exit $?" >> $OUT/${TAG}_dist_tests.log
grep -E "PASSED|FAILED|SKIPPED|passed|failed" $OUT/${TAG}_dist_tests.log | cut -c1-160 | tail -12
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
run bench_n${N} PHDSLAM_MBOX=1 "--steps 10 --warmup 3" 29521
run bench_n${N}_nccl_stats PHDSLAM_MBOX=0 "--steps 10 --warmup 3" 29522
run bench_n${N}_cphd PHDSLAM_MBOX=1 "--workload synthetic_131072x128x50_cphd --steps 5 --warmup 3" 29523
run bench_n${N}_strong16m PHDSLAM_MBOX=1 "--workload synthetic_16777216x128x100_phd --steps 3 --warmup 2" 29524
