#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job38.log
: > $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_seghead_gpu.py -x -q -k "gemm or seghead or conv" --timeout=300 -p no:cacheprovider 2>&1 | tail -n 5 >> $OUT
timeout 300 python tools/kbench.py --only gemm --tag kbench_epi4 >> $OUT 2>&1
cat $OUT
