#!/bin/bash
# round 2, ONE 8-GPU call: scaling of the pretraining step (ViT-Small b256, cfg 3), ViT-Base b128 (cfg 4), the NCCL timeline at
# N = 8, A/B of the communication knobs, and BASELINE.md B1 (the reference's own train() on CUDA) at N = 8.
mkdir -p gpurun_out
OUT=gpurun_out/r2_n8.log
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { # name, port, extra env..., -- bench args
  name=$1; port=$2; shift 2
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 $TR --master-port $port bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - "$name" >> $OUT <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/{n}.json"))
    print(n, "images/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
          "clock", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(n, "FAILED", e)
PY
}
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-stock-baseline > gpurun_out/bench_r02_n1_samebox.json 2> gpurun_out/n1.err
python - >> $OUT <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_n1_samebox.json")); print("n1_samebox images/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 2))
PY
run bench_r02_n8 29531 A=1 -- --steps 30 --warmup 5
run bench_r02_n8_bf16hook 29532 CCD_DDP_BF16_HOOK=1 -- --steps 30 --warmup 5 --no-e2e
run bench_r02_n8_bcast 29533 CCD_DDP_BROADCAST_BUFFERS=1 -- --steps 30 --warmup 5 --no-e2e
run bench_r02_vit_base_b128_n8 29534 A=1 -- --arch vit_base --batch 128 --steps 30 --warmup 5
timeout 300 $TR --master-port 29535 tools/trace_step.py --tag r02_n8 --steps 2 >> $OUT 2>gpurun_out/trace_n8.err
timeout 600 $TR --master-port 29536 bench.py --gpus 8 --impl stock-cuda --steps 6 --warmup 3 > gpurun_out/stock_cuda_r02_n8.json 2> gpurun_out/stock_n8.err
cut -c1-700 gpurun_out/stock_cuda_r02_n8.json >> $OUT
tail -2 gpurun_out/stock_n8.err >> $OUT
cat $OUT
