"""GPU timeline of the pretraining step (CUPTI through torch.profiler; nsys is not in the image).

Runs warm-up steps, profiles `--steps` steps, then reduces the kernel timeline ON THE BOX to a small JSON:
GPU busy / idle per step, the largest idle gaps with the kernels on either side, per-kernel totals.  Works under
torchrun (rank 0 writes).  Output: gpurun_out/trace_<tag>.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ccd_b200 import synthetic as S
from ccd_b200.trainer import FinetuneStep, PretrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--arch", default="vit_small")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--tag", default="n1")
ap.add_argument("--host", action="store_true", help="feed from pinned host memory + loss read (the e2e path)")
ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune"])
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
if a.workload == "finetune":
    t = FinetuneStep(arch=a.arch, batch_per_gpu=a.batch, device=dev, ddp=world > 1)
    t.model.train()
    batch = (torch.randn(a.batch, 3, 32, 128, generator=torch.Generator().manual_seed(1234 + rank)), S.make_targets(a.batch, seed=1234 + rank))
else:
    t = PretrainStep(arch=a.arch, batch_per_gpu=a.batch, device=dev, ddp=world > 1)
    t.student.train()
    batch = S.make_batch(a.batch, seed=1234 + rank)
batch = tuple(v.pin_memory() for v in batch) if a.host else tuple(v.to(dev) for v in batch)
for _ in range(4):
    t.step(*batch, sync_loss=a.host)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.steps):
        t.step(*batch, sync_loss=a.host)
    torch.cuda.synchronize()

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/trace_{a.tag}_raw.json"
    prof.export_chrome_trace(path)
    with open(path) as f:
        tr = json.load(f)
    ev = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    busy, cur_end, gaps = 0.0, t0, []
    per = {}
    prev = None
    for e in ev:
        s, d = e["ts"], e["dur"]
        nm = e["name"][:70]
        k = per.setdefault(nm, [0.0, 0])
        k[0] += d; k[1] += 1
        if s > cur_end:
            gaps.append((s - cur_end, prev, nm, cur_end - t0))
            busy += d
            cur_end = s + d
        else:
            new_end = max(cur_end, s + d)
            busy += new_end - cur_end
            cur_end = new_end
        prev = nm
    gaps.sort(key=lambda g: -g[0])
    span = t1 - t0
    # NCCL kernels run on their own stream: how much of their time is hidden under this rank's compute kernels, and how much
    # is exposed (nothing else running on the GPU)?
    comp = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if not e["name"].startswith("ncclDevKernel"))
    merged = []
    for a0, a1 in comp:
        if merged and a0 <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], a1)
        else:
            merged.append([a0, a1])
    import bisect
    starts = [m[0] for m in merged]
    nccl_total = nccl_exposed = 0.0
    nccl_by = {}
    for e in ev:
        if not e["name"].startswith("ncclDevKernel"):
            continue
        a0, a1 = e["ts"], e["ts"] + e["dur"]
        nccl_total += a1 - a0
        cov = 0.0
        i = max(0, bisect.bisect_right(starts, a0) - 1)
        while i < len(merged) and merged[i][0] < a1:
            cov += max(0.0, min(a1, merged[i][1]) - max(a0, merged[i][0]))
            i += 1
        nccl_exposed += (a1 - a0) - cov
        k = nccl_by.setdefault(e["name"][:48], [0.0, 0.0, 0])
        k[0] += a1 - a0; k[1] += (a1 - a0) - cov; k[2] += 1
    hist = {"<5us": 0, "5-20us": 0, "20-100us": 0, ">100us": 0}
    hsum = dict.fromkeys(hist, 0.0)
    for g in gaps:
        b = "<5us" if g[0] < 5 else "5-20us" if g[0] < 20 else "20-100us" if g[0] < 100 else ">100us"
        hist[b] += 1; hsum[b] += g[0]
    out = {"tag": a.tag, "world": world, "steps": a.steps, "span_ms_per_step": span / 1e3 / a.steps,
           "gpu_busy_ms_per_step": busy / 1e3 / a.steps, "gpu_idle_ms_per_step": (span - busy) / 1e3 / a.steps,
           "n_gpu_events_per_step": len(ev) / a.steps, "gap_histogram_count": hist,
           "gap_histogram_ms_per_step": {k: v / 1e3 / a.steps for k, v in hsum.items()},
           "nccl_ms_per_step": nccl_total / 1e3 / a.steps, "nccl_exposed_ms_per_step": nccl_exposed / 1e3 / a.steps,
           "nccl_kernels": {k: {"ms_per_step": round(v[0] / 1e3 / a.steps, 3), "exposed_ms_per_step": round(v[1] / 1e3 / a.steps, 3),
                                "launches_per_step": v[2] / a.steps} for k, v in nccl_by.items()},
           "largest_gaps": [{"us": round(g[0], 1), "after": g[1], "before": g[2], "at_ms": round(g[3] / 1e3, 2)} for g in gaps[:40]],
           "kernels": sorted(([k, round(v[0] / 1e3 / a.steps, 3), v[1] / a.steps] for k, v in per.items()), key=lambda r: -r[1])[:60]}
    with open(f"gpurun_out/trace_{a.tag}.json", "w") as f:
        json.dump(out, f, indent=1)
    os.remove(path)
    print(json.dumps({k: out[k] for k in ("span_ms_per_step", "gpu_busy_ms_per_step", "gpu_idle_ms_per_step", "gap_histogram_ms_per_step",
                                          "nccl_ms_per_step", "nccl_exposed_ms_per_step", "nccl_kernels")}))
if world > 1:
    dist.destroy_process_group()
