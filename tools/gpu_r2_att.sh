#!/bin/bash
# attention kernels: parity tests, low-overhead timeline of the backward, per-kernel timings
mkdir -p gpurun_out
OUT=gpurun_out/r2_att.log
: > $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "mhsa or kmeans" -p no:cacheprovider 2>&1 | tail -n 8 >> $OUT
CCD_LIB=ccd_b200/libccd_b200_atttrace.so timeout 120 python tools/att_trace.py >> $OUT 2>&1
timeout 300 python tools/kbench.py --only mhsa --tag kbench_r2_att >> $OUT 2>&1
timeout 600 python -m pytest tests/test_pretrain_parity_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -n 8 >> $OUT
cat $OUT
