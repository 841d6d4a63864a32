#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job35.log
: > $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "mhsa" --timeout=300 -p no:cacheprovider 2>&1 | tail -n 6 >> $OUT
for i in 1 2; do timeout 300 python tools/kbench.py --only mhsa --tag kbench_mhsa5 >> $OUT 2>&1; done
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 6 >> $OUT
cat $OUT
