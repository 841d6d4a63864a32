#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job39.log
: > $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_seghead_gpu.py -x -q -k "gemm or seghead or conv" --timeout=300 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
timeout 300 python tools/kbench.py --only gemm --tag kbench_direct >> $OUT 2>&1
CCD_GEMM_EPILOGUE=0 timeout 300 python tools/kbench.py --only gemm --tag kbench_direct_off >> $OUT 2>&1
cat $OUT
