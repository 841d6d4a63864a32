#!/bin/bash
mkdir -p gpurun_out
for t in test_conv3x3_fwd_dgrad_wgrad test_conv_transpose_fwd_dgrad_wgrad test_batchnorm_relu_fwd_bwd test_seghead_against_oracle; do
  echo "=== $t"; timeout 300 python -m pytest tests/test_seghead_gpu.py -q -k "$t" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 25
done
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 8
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile8.json > gpurun_out/bench8.json 2> gpurun_out/bench8.err
cat gpurun_out/bench8.json | cut -c1-300; tail -3 gpurun_out/bench8.err
