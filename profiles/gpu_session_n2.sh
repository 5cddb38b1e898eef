#!/bin/bash
# 2-GPU session: sharding parity test, then the bench at N=1 and N=2 on the same box (weak scaling).
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > $OUT/${TAG}_dist_tests.log 2>&1
tail -3 $OUT/${TAG}_dist_tests.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
cat $OUT/${TAG}_bench_n1.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
cat $OUT/${TAG}_bench_n2.json | cut -c1-1200; tail -3 $OUT/${TAG}_bench_n2.err
# global resampling every step on the streaming shape (configs[4] per-GPU shape): migration traffic
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload synthetic_262144x128x100_phd --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_stream_n2.json 2> $OUT/${TAG}_bench_stream_n2.err
cat $OUT/${TAG}_bench_stream_n2.json | cut -c1-1200; tail -3 $OUT/${TAG}_bench_stream_n2.err
