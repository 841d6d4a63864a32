#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job43.log
: > $OUT
timeout 900 python -m pytest tests/test_finetune_gpu.py -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 8 >> $OUT
timeout 600 python bench.py --workload finetune --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ft_n1.json 2> gpurun_out/bench_ft_n1.err
cut -c1-330 gpurun_out/bench_ft_n1.json >> $OUT; tail -5 gpurun_out/bench_ft_n1.err >> $OUT
timeout 300 python tools/trace_step.py --workload finetune --batch 512 --steps 2 --tag ft2 >> $OUT 2>&1
rm -f gpurun_out/trace_ft2_raw.json
cat $OUT
