#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job40.log
: > $OUT
timeout 300 python tools/kbench.py --only gemm,mhsa --tag kbench_suspend >> $OUT 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 6 >> $OUT
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-330 gpurun_out/bench_n1.json >> $OUT
cat $OUT
