#!/bin/bash
# round 2, 2-GPU session: sharding parity tests with the mailbox exchange and the one-launch push (and the NCCL fallbacks),
# the shim replay, then N=2 bench lines: headline shape (mailbox / NCCL statistics), strong-scaling path at 2 M particles
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_shim_gpu.py -m gpu -v -s --timeout 240 > $OUT/${TAG}_dist_tests.log 2>&1
echo "exit $?" >> $OUT/${TAG}_dist_tests.log
grep -E "PASSED|FAILED|SKIPPED|passed|failed" $OUT/${TAG}_dist_tests.log | cut -c1-200 | tail -16
PHDSLAM_MBOX=0 timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 240 -k "two_gpu" > $OUT/${TAG}_dist_tests_nccl_stats.log 2>&1; tail -1 $OUT/${TAG}_dist_tests_nccl_stats.log
run() { # name, env, args, port
  timeout 600 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 $3 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  python - <<PY
import json
try:
    l=json.loads(open("$OUT/${TAG}_$1.json").read().strip().split("\n")[-1])
    print("$1", round(l["value"]/1e9,2), "G upd/s", round(l["ms_per_step"],3), "ms", {k:round(v,3) for k,v in l["phase_ms"].items()}, l.get("exchange_check"), l["exchange"]["mode"][:20])
except Exception as e:
    print("$1 FAILED", e); print(open("$OUT/${TAG}_$1.err").read()[-1500:])
PY
}
run bench_n2_mbox PHDSLAM_MBOX=1 "--steps 10 --warmup 3" 29511
run bench_n2_ncclstats PHDSLAM_MBOX=0 "--steps 10 --warmup 3" 29512
run bench_n2_resample PHDSLAM_MBOX=1 "--workload synthetic_262144x128x100_phd --steps 3 --warmup 2" 29513
run bench_n2_strong2m PHDSLAM_MBOX=1 "--workload synthetic_2097152x128x100_phd --steps 3 --warmup 2" 29514
