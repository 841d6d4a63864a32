#!/bin/bash
mkdir -p gpurun_out
for v in 1; do
for t in test_gemm_kmajor_bf16_bias test_gemm_majorness_f32 test_gemm_splitk_atomic test_gemm_epilogues; do
  CCD_GEMM_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "$t" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 12
done; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "test_mhsa_fwd or test_layernorm" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 12
timeout 600 python -m pytest tests/test_pretrain_parity_gpu.py -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -n 8
for mv in 0 2; do
CCD_MHSA_FWD_VARIANT=$mv timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/bench_profile3_m$mv.json > gpurun_out/bench3_m$mv.json 2> gpurun_out/bench3.err
cat gpurun_out/bench3_m$mv.json | cut -c1-300; tail -3 gpurun_out/bench3.err
done
CCD_GEMM_VARIANT=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench3_g0.json 2> gpurun_out/bench3.err
cat gpurun_out/bench3_g0.json | cut -c1-300
