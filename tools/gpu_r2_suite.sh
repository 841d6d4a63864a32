#!/bin/bash
# round 2: full GPU test suite + default bench line (with the embedded stock-CUDA baseline)
mkdir -p gpurun_out
OUT=gpurun_out/r2_suite.log
: > $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -n 30 >> $OUT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err
cut -c1-300 gpurun_out/bench_r02b_n1.json >> $OUT
tail -3 gpurun_out/bench_r02b_n1.err >> $OUT
cat $OUT
