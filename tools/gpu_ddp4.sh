#!/bin/bash
# pretrain bench on 4 GPUs of one box (DDP over NCCL/NVLink)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r01f_n4.json 2> gpurun_out/bench_n4.err
cut -c1-330 gpurun_out/bench_r01f_n4.json; grep -v Warning gpurun_out/bench_n4.err | tail -3
