#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python profiles/text_configs.py --impl ours > $OUT/it_text.jsonl 2> $OUT/it_text.err; cat $OUT/it_text.jsonl | cut -c1-900
