#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job23.log
: > $OUT
timeout 900 python -m pytest tests/test_syncbn_2rank_gpu.py tests/test_seghead_gpu.py -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 25 >> $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cut -c1-200 gpurun_out/bench_n2.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n2.json >> $OUT
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-200 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT;  grep -o '"roofline": {[^}]*}' gpurun_out/bench_n1.json | cut -c1-300 >> $OUT
cat $OUT
