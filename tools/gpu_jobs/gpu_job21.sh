#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_ddp.py > gpurun_out/prof_n1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/profile_ddp.py > gpurun_out/prof_n2.log 2>&1
grep -v Warning gpurun_out/prof_n1.log | tail -34; grep -v Warning gpurun_out/prof_n2.log | tail -34
