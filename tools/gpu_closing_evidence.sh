#!/bin/bash
# round-1 closing evidence: full GPU test suite, bench lines (pretrain, finetune, ViT-Base cfg4), ncu launch list of the
# finetune step + full captures of the decoder attention kernels, ViT-Base attention captures
mkdir -p gpurun_out
OUT=gpurun_out/job47.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 4 >> $OUT
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01f_n1.json 2> gpurun_out/bench_r01f_n1.err
cut -c1-300 gpurun_out/bench_r01f_n1.json >> $OUT
timeout 600 python bench.py --workload finetune --steps 10 --warmup 3 > gpurun_out/bench_r01f_finetune_n1.json 2> gpurun_out/bench_r01f_finetune_n1.err
cut -c1-300 gpurun_out/bench_r01f_finetune_n1.json >> $OUT
timeout 600 python bench.py --arch vit_base --batch 128 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01f_vit_base_b128.json 2> gpurun_out/bench_base.err
cut -c1-300 gpurun_out/bench_r01f_vit_base_b128.json >> $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f_finetune.csv \
    --profile-from-start off python tools/profile_step.py --workload finetune --batch 512 --mode list > gpurun_out/ncu_list_ft.log 2>&1
tail -1 gpurun_out/ncu_list_ft.log >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_attn -c 4 --profile-from-start off \
    -o gpurun_out/full_dec_attn_r01f_finetune -f python tools/profile_step.py --workload finetune --batch 512 --mode full > gpurun_out/ncu_dec.log 2>&1
tail -1 gpurun_out/ncu_dec.log >> $OUT
for k in mhsa_fwd_persistent_kernel mhsa_bwd_pipelined_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 --profile-from-start off \
      -o gpurun_out/full_${k}_r01f_base -f python tools/profile_step.py --batch 128 --arch vit_base --mode full > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log >> $OUT
done
cat $OUT
