#!/bin/bash
# configs[4] at N = 2 with the frozen kernels (the N = 4 and N = 8 lines come from r2_session_n8_final.sh)
TAG=${1:-r2s}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload synthetic_16777216x128x100_phd --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_n2_strong16m.json 2> $OUT/${TAG}_bench_n2_strong16m.err
tail -c 1500 $OUT/${TAG}_bench_n2_strong16m.json
