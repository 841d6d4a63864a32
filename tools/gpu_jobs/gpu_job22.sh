#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name: $(python -c "import json;d=json.load(open('gpurun_out/ab_$name.json'));print(d['ms_per_step'], d['value'])")"
}
run default A=1
run noevents CCD_BENCH_NO_EVENTS=1
run nosampler CCD_BENCH_NO_SAMPLER=1
run neither CCD_BENCH_NO_EVENTS=1 CCD_BENCH_NO_SAMPLER=1
run bucketview CCD_DDP_BUCKET_VIEW=1 CCD_BENCH_NO_EVENTS=1 CCD_BENCH_NO_SAMPLER=1
run omp8 OMP_NUM_THREADS=8 CCD_BENCH_NO_EVENTS=1 CCD_BENCH_NO_SAMPLER=1
