"""One CCD pretraining step, as the reference's train.py:221-275 drives it, on the drop-in modules.

This is caller-side glue (what a user's train.py does around Dino.model / Dino.loss); it exists so that bench.py,
smoke() and the tests run exactly the step BASELINE.json's metric is quoted on:
  H2D (optional) -> student fwd -> teacher fwd -> GT-mask warp -> DINOLoss -> backward (+DDP all-reduce)
  -> per-parameter clip -> cancel last-layer grads (epoch < freeze_last_layer) -> AdamW -> teacher EMA.
"""
import math
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .train_utils import (TeacherEMA, cancel_gradients_last_layer, clip_gradients, cosine_iter_scheduler,
                          get_params_groups, has_batchnorms)


class PretrainStep:
    def __init__(self, arch="vit_small", out_dim=65536, batch_per_gpu=256, drop_path_rate=0.1, norm_last_layer=False,
                 lr=0.0005, weight_decay=0.04, weight_decay_end=0.4, clip_grad=3.0, freeze_last_layer=1,
                 momentum_teacher=0.9995, total_iters=100000, device="cuda", ddp=False, seed=0):
        from Dino.loss.Dino_loss import DINOLoss
        from Dino.model.dino_vision import ABIDINOModel
        from Dino.modules import vision_transformer as vits
        from Dino.modules.segmentor import SegHead
        torch.manual_seed(seed)
        self.device = torch.device(device)
        student_bb = vits.__dict__[arch](patch_size=4, drop_path_rate=drop_path_rate)        # train.py:63-67
        teacher_bb = vits.__dict__[arch](patch_size=4)
        E = student_bb.embed_dim
        student = ABIDINOModel(student_bb, SegHead(in_channels=E, mla_channels=128, mlahead_channels=64, num_classes=2),
                               vits.DINOHead(E, out_dim, use_bn=False, norm_last_layer=norm_last_layer))   # :83-86
        teacher = ABIDINOModel(teacher_bb, None, vits.DINOHead(E, out_dim, False))                            # :87-90
        student, teacher = student.to(self.device), teacher.to(self.device)
        self.world = dist.get_world_size() if (ddp and dist.is_initialized()) else 1
        self.student_module, self.teacher_module = student, teacher
        if ddp and dist.is_initialized():
            if has_batchnorms(student):                                                                       # :96-101
                student = nn.SyncBatchNorm.convert_sync_batchnorm(student)
                self.student_module = student
            # train.py:106 uses find_unused_parameters=True (cls_token and segmentation.conv_mla.* never get gradients);
            # static_graph=True gives the same semantics without re-walking the autograd graph every iteration
            # broadcast_buffers: train.py keeps DDP's default (rank 0's BatchNorm running statistics are re-broadcast at every
            # forward); under SyncBatchNorm every rank already holds identical running statistics, so the broadcast -- one more
            # NCCL kernel at the head of every step -- changes nothing and is switched off here (CCD_DDP_BROADCAST_BUFFERS=1 restores it)
            student = nn.parallel.DistributedDataParallel(student, device_ids=[self.device.index],
                                                          find_unused_parameters=True,
                                                          broadcast_buffers=os.environ.get("CCD_DDP_BROADCAST_BUFFERS", "0") == "1",
                                                          gradient_as_bucket_view=os.environ.get("CCD_DDP_BUCKET_VIEW", "1") == "1",
                                                          static_graph=os.environ.get("CCD_DDP_STATIC", "1") == "1")
            if os.environ.get("CCD_DDP_BF16_HOOK", "0") == "1":
                # optional: gradients travel as bf16 (halves the 176.9 MiB all-reduce); off by default -- the reference reduces fp32
                from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                student.register_comm_hook(None, default_hooks.bf16_compress_hook)
        self.student, self.teacher = student, teacher
        self._loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        self._loss_event = torch.cuda.Event()
        teacher.backbone.load_state_dict(self.student_module.backbone.state_dict())                           # :109-110
        teacher.head.load_state_dict(self.student_module.head.state_dict())
        for p in teacher.parameters():
            p.requires_grad = False
        self.loss = DINOLoss(out_dim, 2, 0.04, 0.04, 0, 101).to(self.device)                                  # :122-129
        # train.py:131-133 builds torch.optim.AdamW(params_groups); optim.AdamW is the same optimizer whose step() also
        # folds in the clip (:249), the EMA loop (:264-272) and the bf16 operand refresh (CCD_FUSED_OPT=0: separate calls)
        self.fused_opt = os.environ.get("CCD_FUSED_OPT", "1") == "1"
        if self.fused_opt:
            from .optim import AdamW
            self.opt = AdamW(get_params_groups(student))
        else:
            self.opt = torch.optim.AdamW(get_params_groups(student), fused=True)
        glob = batch_per_gpu * self.world
        self.lr_sched = cosine_iter_scheduler(lr * glob / 256., 1e-6, total_iters,
                                              warmup_iters=min(total_iters // 10, max(1, int(10 * 1000000 / glob))))
        self.wd_sched = cosine_iter_scheduler(weight_decay, weight_decay_end, total_iters)
        self.mom_sched = cosine_iter_scheduler(momentum_teacher, 1, total_iters)
        self.clip_grad, self.freeze_last_layer = clip_grad, freeze_last_layer
        self.ema = TeacherEMA(self.student_module, teacher)
        self.iteration = 0
        self.batch_per_gpu = batch_per_gpu

    def step(self, image_tensors, masks, metrics, epoch=0, sync_loss=True):
        it = self.iteration
        image_tensors = image_tensors.to(self.device, non_blocking=True)                                     # train.py:221-222
        masks = masks.to(self.device, non_blocking=True)
        metrics = metrics.to(self.device, non_blocking=True).float()
        for i, g in enumerate(self.opt.param_groups):                                                         # :224-227
            g["lr"] = float(self.lr_sched[it])
            if i == 0:
                g["weight_decay"] = float(self.wd_sched[it])
        so = self.student(image_tensors, metrics, masks, epoch, clusters=None)                                # :232
        to = self.teacher(image_tensors, metrics, None, None, clusters=so["zero"], index=so["index"])         # :233
        masks_image = ops.warp_mask(masks.contiguous().float(), metrics.contiguous())                         # :234-236
        so["gt"] = [masks, masks_image]
        loss = self.loss(so, to, epoch)                                                                       # :238
        if sync_loss:
            # train.py:239-241 reads loss.item() here and exits on a non-finite value.  The read is issued here but
            # awaited only after the rest of the step has been queued, so the GPU never idles on the host round trip;
            # the process stops on the same iteration either way.
            self._loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            self._loss_event.record()
        self.opt.zero_grad(set_to_none=True)                                                                  # :244
        loss.backward()                                                                                       # :247
        if self.fused_opt:
            cancel_gradients_last_layer(epoch, self.student, self.freeze_last_layer)                          # :250
            self.opt.step(clip_grad=self.clip_grad, ema=self.ema, ema_momentum=float(self.mom_sched[it]))     # :249,252,264-272
        else:
            if self.clip_grad:
                clip_gradients(self.student, self.clip_grad)                                                  # :249
            cancel_gradients_last_layer(epoch, self.student, self.freeze_last_layer)                          # :250
            self.opt.step()                                                                                   # :252
            self.ema.step(float(self.mom_sched[it]))                                                          # :264-272
        self.iteration += 1
        if sync_loss:
            self._loss_event.synchronize()
            if not math.isfinite(float(self._loss_host[0])):
                raise FloatingPointError(f"Loss is {float(self._loss_host[0])}, stopping training")
        return loss


class FinetuneStep:
    """One recognition fine-tuning step as the reference's train_finetune.py:262-290 drives it (BASELINE config 5):
    H2D (optional) -> DINO_Finetune.forward_train (ViT encoder -> Mlp -> NRTR decoder -> TFLoss) -> loss.mean() -> zero_grad
    -> backward (+DDP all-reduce; the reference uses single-process DataParallel) -> [clip_grad_norm_ if configured: the shipped
    configs set clip_grad: ~] -> AdamW with the cosine learning-rate schedule."""

    def __init__(self, arch="vit_small", batch_per_gpu=512, drop_path_rate=0.1, lr=0.0005, weight_decay=0.05, total_iters=100000,
                 device="cuda", ddp=False, seed=0):
        from Dino.model.dino_vision import DINO_Finetune
        from . import synthetic as S
        from .optim import AdamW
        torch.manual_seed(seed)
        self.device = torch.device(device)
        model = DINO_Finetune(S.finetune_config(arch, drop_path_rate)).to(self.device)       # train_finetune.py:185-189
        self.module = model
        self.world = dist.get_world_size() if (ddp and dist.is_initialized()) else 1
        if ddp and dist.is_initialized():
            # backbone.cls_token and backbone.norm_seg.* never receive gradients on this path
            model = nn.parallel.DistributedDataParallel(model, device_ids=[self.device.index], find_unused_parameters=True,
                                                        gradient_as_bucket_view=True, static_graph=True)
        self.model = model
        # train_finetune.py:222-228: AdamW over get_params_groups-style groups (biases / norms without weight decay)
        self.opt = AdamW(get_params_groups(model), lr=lr, weight_decay=weight_decay)
        self.lr_sched = cosine_iter_scheduler(lr, 1e-6, total_iters, warmup_iters=min(total_iters // 10, 1000))
        self._loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        self._loss_event = torch.cuda.Event()
        self.iteration = 0
        self.batch_per_gpu = batch_per_gpu

    def step(self, image_tensors, label_tensors, sync_loss=True):
        it = self.iteration
        image_tensors = image_tensors.to(self.device, non_blocking=True)                     # :277-278
        label_tensors = label_tensors.to(self.device, non_blocking=True)
        for g in self.opt.param_groups:                                                      # :280-281
            g["lr"] = float(self.lr_sched[it])
        losses, _ = self.model(image_tensors, label_tensors, return_loss=True)               # :283
        loss = losses.mean()                                                                 # :284
        if sync_loss:
            self._loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            self._loss_event.record()
        self.opt.zero_grad(set_to_none=True)                                                 # :285
        loss.backward()                                                                      # :286
        self.opt.step()                                                                      # :289
        self.iteration += 1
        if sync_loss:
            self._loss_event.synchronize()
            if not math.isfinite(float(self._loss_host[0])):
                raise FloatingPointError(f"Loss is {float(self._loss_host[0])}, stopping training")
        return loss
