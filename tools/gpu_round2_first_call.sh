#!/bin/bash
# Suggested FIRST gpurun call of the next round (about 2 minutes of box time):
#   1. timeline of the attention backward (build the variant in the CPU container first:
#        python tools/build_variants.py atttrace:CCD_ATT_TRACE=1 )
#      and of the GEMM's roles per tile ( python tools/build_variants.py gemmtrace:CCD_GEMM_TRACE=1 )
#   2. per-kernel timings of every GEMM shape + attention (the table of DESIGN.md section 4)
#   3. the default bench line
mkdir -p gpurun_out
OUT=gpurun_out/round2_first.log
: > $OUT
if [ -f ccd_b200/libccd_b200_atttrace.so ]; then
  CCD_LIB=ccd_b200/libccd_b200_atttrace.so timeout 120 python tools/att_trace.py >> $OUT 2>&1
else
  echo "no ccd_b200/libccd_b200_atttrace.so: run tools/build_variants.py atttrace:CCD_ATT_TRACE=1 first" >> $OUT
fi
if [ -f ccd_b200/libccd_b200_gemmtrace.so ]; then
  CCD_LIB=ccd_b200/libccd_b200_gemmtrace.so timeout 200 python tools/gemm_trace.py >> $OUT 2>&1
fi
timeout 300 python tools/kbench.py --tag kbench_round2 >> $OUT 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_round2_first.json 2>> $OUT
cut -c1-260 gpurun_out/bench_round2_first.json >> $OUT
cat $OUT
