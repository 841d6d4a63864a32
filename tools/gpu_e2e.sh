#!/bin/bash
# end-to-end checks on the GPU box: smoke, parity tests, short bench
mkdir -p gpurun_out
OUT=gpurun_out/e2e.log
: > $OUT
echo "=== smoke" >> $OUT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT 2>&1
echo "=== parity" >> $OUT
timeout 900 python -m pytest tests/test_pretrain_parity_gpu.py -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 60 >> $OUT
echo "=== bench" >> $OUT
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --cpu-batch 8 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.json 2>> $OUT
cat gpurun_out/bench.json >> $OUT
tail -c 6000 $OUT
