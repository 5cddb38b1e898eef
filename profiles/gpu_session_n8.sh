#!/bin/bash
# 8-GPU session: the headline workload (weak scaling, 65 536 particles per GPU) and BASELINE configs[3] at its literal size
# (CPHD, 1 048 576 particles x 128 x 50 over 8 GPUs = 131 072 per GPU)
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
run() { # name, args, port
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 $2 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"phase_ms": {[^}]*}\|"exchange": {[^}]*}' $OUT/${TAG}_$1.json | head -4 | cut -c1-500; tail -2 $OUT/${TAG}_$1.err | cut -c1-300
}
run bench_n8 "--steps 10 --warmup 3" 29521
run bench_cphd_n8 "--workload synthetic_131072x128x50_cphd --steps 5 --warmup 3" 29522
