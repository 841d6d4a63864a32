#!/bin/bash
# Runs every GPU kernel test in its own process (a trapped kernel kills the CUDA context of its process only).
mkdir -p gpurun_out
OUT=gpurun_out/kernel_suite.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $OUT 2>&1
for t in test_gemm_kmajor_bf16_bias test_gemm_majorness_f32 test_gemm_splitk_atomic test_gemm_epilogues \
         test_mhsa_fwd test_mhsa_bwd test_layernorm_fwd_bwd test_colsum_and_norms test_multi_tensor_cast_ema \
         test_dino_ce test_seg_ce test_center_update test_ccl_against_reference_golden test_ccl_random_vs_oracle \
         test_ccl_from_seg_logits test_warp_and_dense test_char_pool test_patch_im2col test_multi_tensor_clip; do
  echo "=== $t" >> $OUT
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "$t" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 25 >> $OUT
done
grep -E "^===|passed|failed|error" $OUT
