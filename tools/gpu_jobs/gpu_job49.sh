#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job49.log
: > $OUT
timeout 600 python -m pytest tests/test_seghead_gpu.py -q --timeout=300 -p no:cacheprovider 2>&1 | grep -v "^E   " | tail -n 25 >> $OUT
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cls.json 2> gpurun_out/bench_cls.err
cut -c1-220 gpurun_out/bench_cls.json >> $OUT; tail -2 gpurun_out/bench_cls.err >> $OUT
timeout 200 python tools/trace_step.py --steps 2 --tag cls >> $OUT 2>&1
rm -f gpurun_out/trace_cls_raw.json
cat $OUT
