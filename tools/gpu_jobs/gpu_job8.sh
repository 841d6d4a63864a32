#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_kernel_suite.sh 2>&1 | tail -25
timeout 600 python -m pytest tests/test_pretrain_parity_gpu.py -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 12
CCD_MHSA_FWD_VARIANT=2 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile5.json > gpurun_out/bench5.json 2> gpurun_out/bench5.err
cat gpurun_out/bench5.json | cut -c1-300; tail -3 gpurun_out/bench5.err
