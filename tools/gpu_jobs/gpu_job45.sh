#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job45.log
: > $OUT
timeout 900 python -m pytest tests/test_finetune_gpu.py -q --timeout=600 -p no:cacheprovider 2>&1 | grep -v "^E  " | tail -n 40 >> $OUT
timeout 600 python bench.py --workload finetune --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ft_n1.json 2> gpurun_out/bench_ft_n1.err
cut -c1-330 gpurun_out/bench_ft_n1.json >> $OUT; tail -3 gpurun_out/bench_ft_n1.err >> $OUT
cat $OUT
