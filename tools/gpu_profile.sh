#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the top kernels (1 GPU only).
mkdir -p gpurun_out
R=${ROUND:-r01}
B=${PBATCH:-256}
# warm-up = 3 steps (skip their launches), then capture every launch of one step
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
    --profile-from-start off python tools/profile_step.py --batch $B --arch ${PARCH:-vit_small} --mode list > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
for k in ${PKERNELS:-gemm_umma_persistent_kernel mhsa_fwd_persistent_kernel mhsa_bwd_kernel layernorm_bwd_kernel dino_ce_fwd_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c ${PCOUNT:-4} --profile-from-start off \
      -o gpurun_out/full_${k}_$R -f python tools/profile_step.py --batch $B --arch ${PARCH:-vit_small} --mode full > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/
