#!/bin/bash
# The N = 8 question: is the step time at N = 8 the slowest GPU's?  Eight INDEPENDENT single-GPU benches, one per
# GPU of the box, at the same time (no communication at all): per-GPU ms/step and clocks.
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i CCD_BENCH_NO_SAMPLER=0 timeout 150 python bench.py --steps 25 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-e2e \
      > gpurun_out/pergpu_$i.json 2> gpurun_out/pergpu_$i.err &
done
wait
python - <<'PY'
import json
out = []
for i in range(8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/pergpu_{i}.json") if l.startswith("{")][-1])
        out.append({"gpu": i, "ms_per_step": round(d["ms_per_step"], 2), "images_per_s": round(d["value"]), "sm_mhz": d["clocks"]["sm_mhz"]})
    except Exception as e:
        out.append({"gpu": i, "error": repr(e)})
print(json.dumps(out))
json.dump(out, open("gpurun_out/pergpu_r02.json", "w"))
PY
