#!/bin/bash
mkdir -p gpurun_out
export CCD_MHSA_FWD_VARIANT=2
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile6.json > gpurun_out/bench6.json 2> gpurun_out/bench6.err
cat gpurun_out/bench6.json | cut -c1-300; tail -3 gpurun_out/bench6.err
