#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job46.log
: > $OUT
timeout 900 python -m pytest tests/test_finetune_gpu.py -q --timeout=600 -p no:cacheprovider -k "greedy or kv_cached" 2>&1 | grep -v "^E  " | tail -n 30 >> $OUT
cat $OUT
