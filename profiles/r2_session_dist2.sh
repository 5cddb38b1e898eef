#!/bin/bash
# round 2, 2-GPU session: the sharded path's parity tests (all 4 cases: NVLink peer window / NCCL ring x PHD / CPHD), full log kept
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -v -s --timeout 240 > $OUT/${TAG}_dist_tests.log 2>&1
echo "exit $?" >> $OUT/${TAG}_dist_tests.log
tail -8 $OUT/${TAG}_dist_tests.log | cut -c1-300
