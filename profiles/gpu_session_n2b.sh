#!/bin/bash
# 2-GPU session: sharding parity in both exchange modes, then the N=2 bench with the NVLink push and with the NCCL ring
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -s --timeout 180 > $OUT/${TAG}_dist_tests.log 2>&1
tail -4 $OUT/${TAG}_dist_tests.log | cut -c1-300
run() { # name, env, workload args
  timeout 300 env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 $3 --no-cpu-baseline > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"phase_ms": {[^}]*}\|"exchange": {[^}]*}' $OUT/${TAG}_$1.json | head -4; tail -2 $OUT/${TAG}_$1.err | cut -c1-300
}
run bench_n2_p2p PHDSLAM_P2P=1 "--steps 10 --warmup 3" 29511
run bench_n2_nccl PHDSLAM_P2P=0 "--steps 10 --warmup 3" 29512
run stream_n2_p2p PHDSLAM_P2P=1 "--workload synthetic_262144x128x100_phd --steps 3 --warmup 2" 29513
run stream_n2_nccl PHDSLAM_P2P=0 "--workload synthetic_262144x128x100_phd --steps 3 --warmup 2" 29514
