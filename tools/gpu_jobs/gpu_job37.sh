#!/bin/bash
# fresh source-level ncu captures of the GEMM epilogue variants (kbench shapes, third launch of each)
mkdir -p gpurun_out
for s in fc1_fwd_gelu_nosave fc2_dgrad_dgelu fc2_fwd_resid; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_persistent -s 2 -c 1 \
      -o gpurun_out/src4_$s -f python tools/kbench.py --only gemm --shape $s --iters 1 --tag tmp > gpurun_out/ncu_$s.log 2>&1
  tail -1 gpurun_out/ncu_$s.log
done
ls -la gpurun_out/*.ncu-rep
