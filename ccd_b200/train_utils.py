"""Caller-side glue of the pretraining step (what train.py:229-272 does around the model/loss calls).

`Dino/modules/utils.py` of the reference is NOT on the hot path of this round: these are small
host-side restatements so that a training step can be driven without the reference tree, plus the multi-tensor
teacher EMA on the sm_100a kernel.
"""

import numpy as np
import torch
import torch.nn as nn

from . import ops


_CLIP_TABLES = {}


def clip_gradients(model, clip):
    """Per-parameter L2 clip (Dino/modules/utils.py:132-141).  On CUDA: two multi-tensor launches, no host sync
    (the reference does one .item() per parameter); returns the squared norms as a device tensor."""
    grads = [p.grad for _, p in model.named_parameters() if p.grad is not None]
    if not grads:
        return []
    if not all(g.is_cuda and g.is_contiguous() and g.dtype == torch.float32 for g in grads):
        raise RuntimeError("ccd_b200.clip_gradients needs contiguous fp32 CUDA gradients (there is no CPU path)")
    table = _CLIP_TABLES.setdefault(id(model), ops.ClipTable())
    return ops.clip_per_parameter_(grads, clip, table).sqrt()


def cancel_gradients_last_layer(epoch, model, freeze_last_layer):
    """Dino/modules/utils.py:144-149."""
    if epoch >= freeze_last_layer:
        return
    for n, p in model.named_parameters():
        if "last_layer" in n:
            p.grad = None


def get_params_groups(model):
    """Dino/modules/utils.py:643-654: biases and 1-D (norm) parameters are not weight-decayed."""
    reg, noreg = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (noreg if name.endswith(".bias") or p.ndim == 1 else reg).append(p)
    return [{"params": reg}, {"params": noreg, "weight_decay": 0.}]


def has_batchnorms(model):
    bn = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm)
    return any(isinstance(m, bn) for m in model.modules())


def cosine_iter_scheduler(base_value, final_value, niter, warmup_iters=0, start_warmup_value=0):
    """Dino/modules/utils.py:200-210."""
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_iters > 0 else np.array([])
    it = np.arange(niter - warmup_iters)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * it / len(it)))
    out = np.concatenate((warm, sched))
    assert len(out) == niter
    return out


class TeacherEMA:
    """train.py:264-272 as ONE multi-tensor kernel: p_t = m p_t + (1-m) p_s over backbone.* and head.* pairs."""

    def __init__(self, student, teacher):
        self.pairs = list(zip(student.backbone.parameters(), teacher.backbone.parameters())) + \
            list(zip(student.head.parameters(), teacher.head.parameters()))
        self.table = ops.ChunkTable()
        # modules holding bf16 GEMM-operand copies of parameters (refreshed by the fused optimizer step, optim.AdamW)
        self.bf16_modules = [m for m in (student.backbone, student.head, teacher.backbone, teacher.head)
                             if hasattr(m, "bf16_copies")]

    @torch.no_grad()
    def step(self, m):
        srcs = [s.detach() for s, _ in self.pairs]
        dsts = [t.detach() for _, t in self.pairs]
        table, n = self.table.get(srcs, dsts, 4)
        ops.multi_tensor(ops.MT_EMA, table, n, float(m), float(1.0 - m))
        torch.autograd.graph.increment_version(dsts)      # raw-pointer in-place update: keep version counters honest
