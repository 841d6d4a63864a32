#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job42.log
: > $OUT
timeout 600 python bench.py --workload finetune --steps 10 --warmup 3 > gpurun_out/bench_ft_n1.json 2> gpurun_out/bench_ft_n1.err
cut -c1-1500 gpurun_out/bench_ft_n1.json >> $OUT; tail -5 gpurun_out/bench_ft_n1.err >> $OUT
timeout 300 python tools/trace_step.py --workload finetune --batch 512 --steps 2 --tag ft >> $OUT 2>&1
rm -f gpurun_out/trace_ft_raw.json
cat $OUT
