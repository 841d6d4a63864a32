#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke(), the default bench line, the finetune bench line
mkdir -p gpurun_out
OUT=gpurun_out/final_check.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 3 >> $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 >> $OUT
timeout 600 python bench.py > gpurun_out/bench_r01g_n1.json 2> gpurun_out/bench_final.err
cut -c1-260 gpurun_out/bench_r01g_n1.json >> $OUT; tail -2 gpurun_out/bench_final.err >> $OUT
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01g_reference.json 2>> gpurun_out/bench_final.err
cut -c1-260 gpurun_out/bench_r01g_reference.json >> $OUT
cat $OUT
