"""Runs warm-up steps, then ONE profiled pretraining step between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccd_b200 import synthetic as S
from ccd_b200.trainer import PretrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--arch", default="vit_small")
ap.add_argument("--mode", default="list")
ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune"])
a = ap.parse_args()
if a.workload == "finetune":
    from ccd_b200.trainer import FinetuneStep
    t = FinetuneStep(arch=a.arch, batch_per_gpu=a.batch, device=torch.device("cuda", 0))
    t.model.train()
    batch = (torch.randn(a.batch, 3, 32, 128, generator=torch.Generator().manual_seed(1234)).cuda(), S.make_targets(a.batch, seed=1234).cuda())
else:
    t = PretrainStep(arch=a.arch, batch_per_gpu=a.batch, device=torch.device("cuda", 0))
    t.student.train()
    batch = tuple(v.cuda() for v in S.make_batch(a.batch, seed=1234))
for _ in range(2):
    t.step(*batch, sync_loss=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t.step(*batch, sync_loss=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step, batch", a.batch)
