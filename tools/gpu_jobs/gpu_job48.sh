#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/job48.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 3 >> $OUT
for v in 1 0 1 0; do
  CCD_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl.err
  echo "PDL=$v $(cut -c1-220 gpurun_out/bench_pdl$v.json | grep -o 'ms_per_step[^,]*')" >> $OUT
done
tail -3 gpurun_out/bench_pdl.err >> $OUT
cat $OUT
