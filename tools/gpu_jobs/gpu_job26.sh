#!/bin/bash
# batched GEMM epilogue + host-stall fixes + fused optimizer: tests, kernel A/B, bench, timeline
mkdir -p gpurun_out
OUT=gpurun_out/job26.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
timeout 600 python tools/kbench.py --only gemm --tag kbench_epi2 >> $OUT 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-260 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
CCD_FUSED_OPT=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>> $OUT | cut -c1-260 >> $OUT
timeout 300 python tools/trace_step.py --tag n1b >> $OUT 2>&1
cat $OUT
