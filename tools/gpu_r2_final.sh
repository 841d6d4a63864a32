#!/bin/bash
# round 2 closing evidence on ONE GPU: full GPU test suite, default bench line (stock-CUDA + CPU reference baselines embedded),
# finetune bench line, gradient-parity table
mkdir -p gpurun_out
OUT=gpurun_out/r2_final.log
: > $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=800 -p no:cacheprovider 2>&1 | tail -n 6 >> $OUT
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02c_n1.json 2> gpurun_out/bench_r02c_n1.err
grep '^{' gpurun_out/bench_r02c_n1.json | cut -c1-260 >> $OUT
timeout 300 python bench.py --workload finetune --steps 20 --warmup 5 > gpurun_out/bench_r02c_finetune_n1.json 2> gpurun_out/bench_ft.err
grep '^{' gpurun_out/bench_r02c_finetune_n1.json | cut -c1-260 >> $OUT
timeout 300 python tests/grad_parity_report.py cfg1_tiny_b4 small_b32_k65536 --md gpurun_out/GRAD_PARITY.md > gpurun_out/grad_parity.log 2>&1
tail -4 gpurun_out/grad_parity.log >> $OUT
cat $OUT
