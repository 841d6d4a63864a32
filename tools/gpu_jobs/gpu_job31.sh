#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_bwd_pipelined -s 2 -c 1 \
    -o gpurun_out/src3_mhsa_bwd -f python tools/kbench.py --only mhsa --iters 1 --tag tmp > gpurun_out/ncu_mb3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_fwd_persistent -s 2 -c 1 \
    -o gpurun_out/src3_mhsa_fwd -f python tools/kbench.py --only mhsa --iters 1 --tag tmp > gpurun_out/ncu_mf3.log 2>&1
tail -2 gpurun_out/ncu_mb3.log gpurun_out/ncu_mf3.log
