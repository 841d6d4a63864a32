#!/bin/bash
# round 2, 2 GPUs of one box: this path at N = 1 and N = 2, BASELINE.md B1 (the reference's own train() + modules on CUDA, DDP +
# SyncBN exactly as train.py:93-110) at N = 2, fine-tuning step at N = 2
mkdir -p gpurun_out
OUT=gpurun_out/r2_n2.log
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-stock-baseline > gpurun_out/bench_r02_n1_samebox2.json 2> gpurun_out/n1b.err
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/n2.err
timeout 420 $TR --master-port 29542 bench.py --gpus 2 --impl stock-cuda --steps 6 --warmup 3 > gpurun_out/stock_cuda_r02_n2.json 2> gpurun_out/stock_n2.err
timeout 300 $TR --master-port 29543 bench.py --gpus 2 --workload finetune --steps 20 --warmup 5 > gpurun_out/bench_r02_finetune_n2.json 2> gpurun_out/ft_n2.err
python - >> $OUT <<'PY'
import json
for n in ("bench_r02_n1_samebox2", "bench_r02_n2", "bench_r02_finetune_n2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, "images/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "clock", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "FAILED", repr(e))
try:
    print(open("gpurun_out/stock_cuda_r02_n2.json").read()[-900:])
except Exception as e:
    print("stock FAILED", e)
PY
tail -3 gpurun_out/stock_n2.err >> $OUT
cat $OUT
