#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job20.log
: > $OUT
echo "=== full gpu suite" >> $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
echo "=== bench" >> $OUT
timeout 900 python bench.py --steps 8 --warmup 3 --cpu-batch 8 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.json 2>> $OUT
cat gpurun_out/bench.json >> $OUT
tail -c 4000 $OUT
echo "=== ddp2"
bash tools/gpu_ddp2.sh
