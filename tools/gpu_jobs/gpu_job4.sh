#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pretrain_parity_gpu.py -q --timeout=600 -p no:cacheprovider 2>&1 | tail -n 40 > gpurun_out/parity.log
tail -5 gpurun_out/parity.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "clip" -p no:cacheprovider 2>&1 | tail -3
bash tools/gpu_profile.sh
