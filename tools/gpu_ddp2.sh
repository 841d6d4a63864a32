#!/bin/bash
# 2-GPU DDP run of the bench (NCCL over NVLink), plus the 1-GPU number on the same box for the scaling ratio
mkdir -p gpurun_out
export CCD_MHSA_FWD_VARIANT=${CCD_MHSA_FWD_VARIANT:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json | cut -c1-400; tail -5 gpurun_out/bench_n2.err
timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json | cut -c1-300
