#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -k "$1" > $OUT/it_tests.log 2>&1; tail -5 $OUT/it_tests.log
