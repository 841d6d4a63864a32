#!/bin/bash
# N=2: SyncBN 2-rank parity test, bench at N=2 and N=1 on the same box, CUPTI timeline of the 2-rank step
mkdir -p gpurun_out
OUT=gpurun_out/job36.log
: > $OUT
timeout 300 python -m pytest tests/test_syncbn_2rank_gpu.py -x -q --timeout=250 -p no:cacheprovider 2>&1 | tail -n 4 >> $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cut -c1-330 gpurun_out/bench_n2.json >> $OUT; tail -3 gpurun_out/bench_n2.err >> $OUT
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-330 gpurun_out/bench_n1.json >> $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    tools/trace_step.py --steps 2 --tag n2 >> $OUT 2>&1
rm -f gpurun_out/trace_n2_raw.json
cat $OUT
