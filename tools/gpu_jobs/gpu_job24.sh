#!/bin/bash
# validation of HEAD after container re-creation: full GPU suite, bench N=1 (+per-shape profile), CUPTI timeline
mkdir -p gpurun_out
OUT=gpurun_out/job24.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --profile-out gpurun_out/bench_profile_r01e.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
timeout 300 python tools/trace_step.py --tag n1 >> $OUT 2>&1
timeout 300 python tools/trace_step.py --tag n1_host --host >> $OUT 2>&1
cat $OUT
