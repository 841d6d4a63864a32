#!/bin/bash
# kernel-level baseline timings + ncu source-level captures of the epilogue-bound GEMMs and the attention kernels
mkdir -p gpurun_out
OUT=gpurun_out/job25.log
: > $OUT
timeout 600 python tools/kbench.py --tag kbench_base >> $OUT 2>&1
for sh in fc1_fwd_gelu_save fc2_dgrad_dgelu qkv_fwd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_persistent -s 2 -c 1 \
      -o gpurun_out/src_${sh} -f python tools/kbench.py --only gemm --shape $sh --iters 1 --tag tmp > gpurun_out/ncu_$sh.log 2>&1
  tail -2 gpurun_out/ncu_$sh.log >> $OUT
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_fwd_persistent -s 2 -c 1 \
    -o gpurun_out/src_mhsa_fwd -f python tools/kbench.py --only mhsa --iters 1 --tag tmp > gpurun_out/ncu_mf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_bwd_kernel -s 2 -c 1 \
    -o gpurun_out/src_mhsa_bwd -f python tools/kbench.py --only mhsa --iters 1 --tag tmp > gpurun_out/ncu_mb.log 2>&1
tail -2 gpurun_out/ncu_mf.log gpurun_out/ncu_mb.log >> $OUT
ls -la gpurun_out >> $OUT
cat $OUT
