#!/bin/bash
# both epilogue warpgroups per tile + bias prefetch + packed tanh GELU: tests, kernel A/B (tanh vs ex2/rcp), bench
mkdir -p gpurun_out
OUT=gpurun_out/job28.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
cat gpurun_out/gelu_accuracy.json >> $OUT; echo >> $OUT
timeout 600 python tools/kbench.py --only gemm --tag kbench_epi3 >> $OUT 2>&1
echo "== notanh" >> $OUT
CCD_LIB=$PWD/ccd_b200/libccd_b200_notanh.so timeout 300 python tools/kbench.py --only gemm --shape fc --tag kbench_notanh 2>&1 | tail -1 >> $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-260 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
timeout 300 python tools/trace_step.py --tag n1c --steps 3 >> $OUT 2>&1
cat $OUT
