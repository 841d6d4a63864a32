#!/bin/bash
# pipelined mhsa fwd + indirect-pointer fused optimizer: tests, A/B of epilogue experiment builds, bench
mkdir -p gpurun_out
OUT=gpurun_out/job27.log
: > $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -p no:cacheprovider 2>&1 | tail -n 15 >> $OUT
timeout 300 python tools/kbench.py --only mhsa --tag kbench_mhsa2 >> $OUT 2>&1
for v in dbg1 dbg2 dbg3; do
  echo "== $v" >> $OUT
  CCD_LIB=$PWD/ccd_b200/libccd_b200_$v.so timeout 300 python tools/kbench.py --only gemm --shape fc --tag kbench_$v 2>&1 | tail -1 >> $OUT
  CCD_LIB=$PWD/ccd_b200/libccd_b200_$v.so timeout 300 python tools/kbench.py --only gemm --shape qkv_fwd --tag kbench_${v}q 2>&1 | tail -1 >> $OUT
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-260 gpurun_out/bench_n1.json >> $OUT; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json >> $OUT; tail -3 gpurun_out/bench_n1.err >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_persistent -s 2 -c 1 \
      -o gpurun_out/src2_fc1 -f python tools/kbench.py --only gemm --shape fc1_fwd_gelu_save --iters 1 --tag tmp > gpurun_out/ncu_fc1b.log 2>&1
cat $OUT
