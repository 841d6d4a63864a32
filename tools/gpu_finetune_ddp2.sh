#!/bin/bash
# finetune workload on 2 GPUs (DDP over NCCL) + 1 GPU on the same box
mkdir -p gpurun_out
OUT=gpurun_out/job44.log
: > $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --workload finetune --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_ft_n2.json 2> gpurun_out/bench_ft_n2.err
cut -c1-330 gpurun_out/bench_ft_n2.json >> $OUT; grep -v Warning gpurun_out/bench_ft_n2.err | tail -3 >> $OUT
timeout 300 python bench.py --workload finetune --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ft_n1.json 2> gpurun_out/bench_ft_n1.err
cut -c1-330 gpurun_out/bench_ft_n1.json >> $OUT
cat $OUT
