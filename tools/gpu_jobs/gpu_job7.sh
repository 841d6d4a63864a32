#!/bin/bash
mkdir -p gpurun_out
for t in test_gemm_kmajor_bf16_bias test_gemm_majorness_f32 test_gemm_splitk_atomic test_gemm_epilogues; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "$t" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 12
done
timeout 600 python -m pytest tests/test_pretrain_parity_gpu.py -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 12
CCD_MHSA_FWD_VARIANT=2 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile4.json > gpurun_out/bench4.json 2> gpurun_out/bench4.err
cat gpurun_out/bench4.json | cut -c1-300; tail -3 gpurun_out/bench4.err
timeout 900 python bench.py --impl stock-eager-cuda --steps 3 --warmup 2 > gpurun_out/stock_eager.json 2> gpurun_out/stock.err
cat gpurun_out/stock_eager.json; tail -3 gpurun_out/stock.err
