#!/bin/bash
# round 2 ncu evidence (1 GPU): launch list of one pretraining step, dram bytes of EVERY GEMM launch of a step (per-shape traffic),
# full captures of the attention kernels at ViT-Small and ViT-Base (BASELINE cfg 4) and of a few GEMM launches.
mkdir -p gpurun_out
R=r02
OUT=gpurun_out/r2_profile.log
: > $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
    --profile-from-start off python tools/profile_step.py --batch 256 --mode list > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log >> $OUT
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_umma \
    --profile-from-start off --csv --log-file gpurun_out/gemm_traffic_$R.csv python tools/gemm_traffic.py --run > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mhsa_(fwd_persistent|bwd_pipelined)' -c 4 --profile-from-start off \
    -o gpurun_out/full_mhsa_$R -f python tools/profile_step.py --batch 256 --mode full > gpurun_out/ncu_mhsa.log 2>&1
tail -1 gpurun_out/ncu_mhsa.log >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mhsa_(fwd_persistent|bwd_pipelined)' -c 4 --profile-from-start off \
    -o gpurun_out/full_mhsa_base_$R -f python tools/profile_step.py --batch 128 --arch vit_base --mode full > gpurun_out/ncu_mhsa_base.log 2>&1
tail -1 gpurun_out/ncu_mhsa_base.log >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_persistent -s 30 -c 12 --profile-from-start off \
    -o gpurun_out/full_gemm_$R -f python tools/profile_step.py --batch 256 --mode full > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log >> $OUT
ls -la gpurun_out/*_$R* >> $OUT
cat $OUT
