"""Phase timing (CUDA events) + torch.profiler kernel table of the pretraining step, with or without DDP.
    python tools/profile_ddp.py                       # 1 GPU
    torchrun --nproc-per-node 2 tools/profile_ddp.py  # DDP"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ccd_b200 import ops, synthetic as S
from ccd_b200.trainer import PretrainStep
from ccd_b200.train_utils import clip_gradients, cancel_gradients_last_layer

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = PretrainStep(arch="vit_small", batch_per_gpu=256, device=torch.device("cuda", lr), ddp=world > 1)
t.student.train()
x, m, th = [v.cuda() for v in S.make_batch(256, seed=1234 + rank)]
for _ in range(4):
    t.step(x, m, th, sync_loss=False)
torch.cuda.synchronize()


def phased():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    k = 0
    def mark():
        nonlocal k
        ev[k].record(); k += 1
    self = t
    mark()
    so = self.student(x, th.float(), m, 0, clusters=None); mark()
    to = self.teacher(x, th.float(), None, None, clusters=so["zero"], index=so["index"]); mark()
    so["gt"] = [m, ops.warp_mask(m.contiguous().float(), th.contiguous().float())]
    loss = self.loss(so, to, 0); mark()
    self.opt.zero_grad(set_to_none=True)
    loss.backward(); mark()
    clip_gradients(self.student, self.clip_grad); mark()
    cancel_gradients_last_layer(0, self.student, self.freeze_last_layer)
    self.opt.step(); mark()
    self.ema.step(0.9995); mark()
    torch.cuda.synchronize()
    names = ["student fwd", "teacher fwd", "loss", "backward", "clip", "adamw", "ema"]
    return {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}

acc = {}
for _ in range(4):
    for n, v in phased().items():
        acc[n] = acc.get(n, 0.0) + v / 4
if rank == 0:
    print(f"world {world} phases (ms):", {n: round(v, 2) for n, v in acc.items()}, "total", round(sum(acc.values()), 2), flush=True)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        t.step(x, m, th, sync_loss=False)
    torch.cuda.synchronize()
if rank == 0:
    rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type is not None]
    ker = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ker.setdefault(e.name[:90], [0.0, 0])
            ker[e.name[:90]][0] += e.device_time_total / 2e3
            ker[e.name[:90]][1] += 1
    tot = sum(v[0] for v in ker.values())
    print(f"world {world}: CUDA busy {tot:.2f} ms/step over {sum(v[1] for v in ker.values()) // 2} launches")
    for n, (ms, c) in sorted(ker.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"  {ms:8.3f} ms {c // 2:5d}  {n}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
