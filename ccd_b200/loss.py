"""DINO-style character distillation loss (drop-in for DINOLoss / SegLoss, Dino/loss/Dino_loss.py:7-143).

Same constructor, `forward(student_output, teacher_output, epoch)`, `last_losses` property, `center` buffer and
`teacher_temp_schedule` as the reference.  The sharpened-softmax cross-entropy (forward + backward), the teacher
column sums for the centre, the centre EMA and the segmentation cross-entropy are sm_100a kernels; the centre
all-reduce rides NCCL through torch.distributed exactly where the reference calls it (Dino_loss.py:139).
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .head import DLOGITS_STASH, head_logits_registered


class DinoCEFn(torch.autograd.Function):
    """L = 1/2 [ mean_r CE(q_0, p_1) + mean_r CE(q_1, p_0) ]   (Dino_loss.py:81-105 with ncrops = 2)."""

    @staticmethod
    def forward(ctx, zs, zt, center, student_temp, teacher_temp):
        zs_c, zt_c = zs.contiguous().float(), zt.detach().contiguous().float()
        loss, stats = ops.dino_ce_fwd(zs_c, zt_c, center.contiguous().float().view(-1), student_temp, teacher_temp)
        # the centre is updated in place right after the forward (Dino_loss.py:104): backward needs the OLD centre
        ctx.save_for_backward(zs_c, zt_c, center.detach().clone().view(-1), stats)
        ctx.temps = (student_temp, teacher_temp)
        # the bf16 hand-off to HeadFn.backward is only valid when zs IS the tensor a ccd_b200 DINOHead produced
        ctx.zs_ptr = zs.data_ptr() if head_logits_registered(zs) else None
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        zs, zt, center, stats = ctx.saved_tensors
        ts, tt = ctx.temps
        dz = ops.dino_ce_bwd(zs, zt, center, stats, g.contiguous().float().view(1), ts, tt)
        if ctx.zs_ptr is None:
            return dz.float(), None, None, None, None          # foreign logits: an ordinary dense gradient
        # fast path: hand the bf16 dlogits to HeadFn.backward (it consumes them as the GEMM A operand); autograd itself
        # only sees a zero-stride placeholder, so no fp32 [2R,K] gradient is ever materialised.
        placeholder = torch.zeros(1, 1, dtype=zs.dtype, device=zs.device)
        DLOGITS_STASH[ctx.zs_ptr] = (dz, placeholder.data_ptr())
        return placeholder.expand(zs.shape), None, None, None, None


class SegCEFn(torch.autograd.Function):
    """CE applied to already-softmaxed probabilities, mean over pixels (Dino_loss.py:63-68,15-26)."""

    @staticmethod
    def forward(ctx, logits, gt):
        lg, gtf = logits.contiguous().float(), gt.contiguous().float()
        loss = ops.seg_ce_fwd(lg, gtf)
        ctx.save_for_backward(lg, gtf)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        lg, gtf = ctx.saved_tensors
        return ops.seg_ce_bwd(lg, gtf, g.contiguous().float().view(1)), None


class SegLoss(nn.Module):
    def __init__(self, loss_seg=False):
        super().__init__()
        self.loss_seg = loss_seg

    def cross_entropy(self, global_text_segs, gts_masks, bool_=True):
        if global_text_segs.shape[-1] != gts_masks.shape[-1]:
            raise NotImplementedError("ccd_b200.SegLoss: prediction and target must have the same resolution")
        if not bool_:
            raise NotImplementedError("ccd_b200.SegLoss: only the reduced (mean) loss is implemented")
        return SegCEFn.apply(global_text_segs, gts_masks)

    def forward(self, seg_mask, gts_masks, bool_=True):
        return self.cross_entropy(seg_mask, gts_masks, bool_)


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9):
        super().__init__()
        if ncrops != 2:
            raise NotImplementedError("the reference model produces exactly 2 views (SURVEY F1); ncrops must be 2")
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = np.concatenate((                                   # Dino_loss.py:46-50
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp))
        self.seg_loss = SegLoss()
        self.losses = {}

    @property
    def last_losses(self):
        return self.losses

    def forward(self, student_output, teacher_output, epoch):
        self.losses = {}
        gt = torch.cat(student_output["gt"])                                            # [masks, masks_image]
        # the reference soft-maxes here and again inside F.cross_entropy; SegCEFn computes exactly that composition
        mask_loss = SegCEFn.apply(student_output["mask"], gt)
        self.losses["mask_loss"] = mask_loss
        zs = student_output["instances_view"]
        zt = teacher_output["instances_view"]
        temp = float(self.teacher_temp_schedule[epoch])
        dino_loss = DinoCEFn.apply(zs, zt, self.center, self.student_temp, temp)
        self.update_center(zt)
        self.losses["Dino_loss"] = dino_loss
        return mask_loss + dino_loss

    @torch.no_grad()
    def update_center(self, teacher_output):
        """Dino_loss.py:133-143: SUM over ranks of the per-rank row sums, divided by local_rows * world_size."""
        distributed = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size() if distributed else 1
        ops.center_update(self.center.view(-1), teacher_output.detach().contiguous().float(), world, self.center_momentum,
                          all_reduce=(dist.all_reduce if distributed else None))
