#!/bin/bash
mkdir -p gpurun_out
# CPU reference arm on this box's host cores (bounded)
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
cat gpurun_out/ref_arm.json | cut -c1-700; tail -4 gpurun_out/ref_arm.err
# other architectures of the reference (configs 1 and 4 of BASELINE.json)
timeout 600 python bench.py --arch vit_base --batch 128 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile_base.json > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
cat gpurun_out/bench_base.json | cut -c1-260; tail -2 gpurun_out/bench_base.err
timeout 600 python bench.py --arch vit_tiny --batch 256 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err
cat gpurun_out/bench_tiny.json | cut -c1-260; tail -2 gpurun_out/bench_tiny.err
# config 4: attention kernels at ViT-Base shapes (2048 problems per layer), full ncu capture
ROUND=r01_base PARCH=vit_base PBATCH=128 PKERNELS="mhsa_fwd_persistent_kernel mhsa_bwd_kernel" PCOUNT=3 bash tools/gpu_profile.sh 2>&1 | tail -3
