#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_seghead_gpu.py -q -k "test_seghead_against_oracle" --timeout=240 -p no:cacheprovider 2>&1 | tail -n 12
ROUND=r01d PKERNELS="none" PCOUNT=1 bash tools/gpu_profile.sh 2>&1 | tail -3
