#!/bin/bash
mkdir -p gpurun_out
for t in test_gemm_kmajor_bf16_bias test_gemm_majorness_f32 test_gemm_splitk_atomic test_gemm_epilogues; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "$t" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 8
done
timeout 600 python tests/grad_parity_report.py cfg1_tiny_b4 > gpurun_out/grad_report_tiny.txt 2>&1; tail -5 gpurun_out/grad_report_tiny.txt
timeout 600 python tests/grad_parity_report.py small_b3 > gpurun_out/grad_report_small.txt 2>&1; tail -3 gpurun_out/grad_report_small.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile2.json > gpurun_out/bench2.json 2> gpurun_out/bench2.err
cat gpurun_out/bench2.json | cut -c1-1500; tail -5 gpurun_out/bench2.err
