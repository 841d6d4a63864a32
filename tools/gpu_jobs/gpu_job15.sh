#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile7.json > gpurun_out/bench7.json 2> gpurun_out/bench7.err
cat gpurun_out/bench7.json | cut -c1-300; tail -3 gpurun_out/bench7.err
ROUND=r01c PKERNELS="gemm_umma_persistent_kernel mhsa_bwd_kernel mhsa_fwd_persistent_kernel layernorm_bwd_kernel" PCOUNT=8 bash tools/gpu_profile.sh 2>&1 | tail -4
