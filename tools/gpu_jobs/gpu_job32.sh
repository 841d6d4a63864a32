#!/bin/bash
# BN streaming kernels rewrite + cp.async staging in the cls convolution kernels
mkdir -p gpurun_out
OUT=gpurun_out/job32.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 8 >> $OUT
timeout 300 python tools/trace_step.py --tag n1d --steps 3 >> $OUT 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-260 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
cat $OUT
