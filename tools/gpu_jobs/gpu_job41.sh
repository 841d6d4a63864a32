#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job41.log
: > $OUT
timeout 900 python -m pytest tests/test_finetune_gpu.py -x -q --timeout=600 -p no:cacheprovider -s 2>&1 | tail -n 40 >> $OUT
cat $OUT
