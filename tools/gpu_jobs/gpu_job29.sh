#!/bin/bash
# pipelined persistent attention backward: parity (both variants), timing, full suite, bench
mkdir -p gpurun_out
OUT=gpurun_out/job29.log
: > $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "mhsa" --timeout=300 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
timeout 300 python tools/kbench.py --only mhsa --tag kbench_mhsa3 >> $OUT 2>&1
CCD_MHSA_BWD_VARIANT=0 timeout 300 python tools/kbench.py --only mhsa --tag kbench_mhsa3_v0 >> $OUT 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 8 >> $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-260 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
cat $OUT
