#!/bin/bash
# seghead cls kernels: unit tests, the whole gpu suite in one process, bench with per-entry-point profile
mkdir -p gpurun_out
OUT=gpurun_out/job19.log
: > $OUT
echo "=== seghead tests" >> $OUT
timeout 600 python -m pytest tests/test_seghead_gpu.py -q --timeout=300 -p no:cacheprovider 2>&1 | tail -n 30 >> $OUT
echo "=== full gpu suite" >> $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
echo "=== bench" >> $OUT
timeout 900 python bench.py --steps 8 --warmup 3 --cpu-batch 8 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.json 2>> $OUT
cat gpurun_out/bench.json >> $OUT
echo "=== launch list" >> $OUT
ROUND=r01e PCOUNT=2 PKERNELS="seg_cls_fwd_kernel seg_cls_wgrad_kernel seg_cls_dgrad_kernel" bash tools/gpu_profile.sh >> $OUT 2>&1 || true
tail -c 5000 $OUT
