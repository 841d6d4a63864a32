"""TEST / BASELINE INFRASTRUCTURE -- one CCD pretraining step computed by the UNMODIFIED reference modules
(/root/reference, or the hash-checked byte-compiled tree oracle/_ref on the GPU box), on any device.

The loop body below is this harness's own (the reference's `train()` needs CUDA, NCCL and LMDB data; oracle/run_ref_train.py
runs that one on a GPU); every quantity is produced by the reference's classes:
    ABIDINOModel / vit_* / SegHead / DINOHead / DINOLoss          Dino/model/dino_vision.py, Dino/modules/*, Dino/loss/Dino_loss.py
    clip_gradients, cancel_gradients_last_layer, get_params_groups  Dino/modules/utils.py:132-149,643-654
in the order of train.py:229-272 (forward student, forward teacher, GT warp, loss, backward, clip, cancel, AdamW, EMA).

Users: bench.py --impl reference (cpu_baseline.kind = "reference", host cores) and tests/ (on the GPU: fp32 eager reference
as the comparator at BASELINE's full sizes).  Never the product path.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ref_import  # noqa: E402

EMBED = {"vit_tiny": 192, "vit_small": 384, "vit_base": 512}


class ReferenceStep:
    def __init__(self, arch="vit_small", out_dim=65536, drop_path_rate=0.0, norm_last_layer=False, device="cpu",
                 student_sd=None, teacher_sd=None, lr=0.0005, weight_decay=0.04, clip_grad=3.0, freeze_last_layer=1,
                 momentum_teacher=0.9995, nepochs=101):
        self.ref = ref = ref_import.load_reference()
        ref_import.ensure_gloo_group()                      # DINOLoss.update_center all-reduces unconditionally (Dino_loss.py:139)
        E = EMBED[arch]
        dev = torch.device(device)
        self.student = ref.dv.ABIDINOModel(getattr(ref.vits, arch)(patch_size=4, drop_path_rate=drop_path_rate),
                                           ref.seg.SegHead(in_channels=E, mla_channels=128, mlahead_channels=64, num_classes=2),
                                           ref.vits.DINOHead(E, out_dim, use_bn=False, norm_last_layer=norm_last_layer))
        self.teacher = ref.dv.ABIDINOModel(getattr(ref.vits, arch)(patch_size=4), None, ref.vits.DINOHead(E, out_dim, False))
        if student_sd is not None:
            self.student.load_state_dict(student_sd)
        if teacher_sd is not None:
            self.teacher.load_state_dict(teacher_sd)
        else:                                               # train.py:109-110
            self.teacher.backbone.load_state_dict(self.student.backbone.state_dict())
            self.teacher.head.load_state_dict(self.student.head.state_dict())
        self.student, self.teacher = self.student.to(dev), self.teacher.to(dev)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.loss = ref.loss.DINOLoss(out_dim, 2, 0.04, 0.04, 0, nepochs).to(dev)
        self.opt = torch.optim.AdamW(ref.utils.get_params_groups(self.student), lr=lr, weight_decay=weight_decay)
        self.clip_grad, self.freeze_last_layer, self.m = clip_grad, freeze_last_layer, momentum_teacher
        self.device = dev

    def forward(self, x, masks, metrics, epoch=0):
        """train.py:229-238: returns (loss, student_output, teacher_output)."""
        x, masks, metrics = x.to(self.device), masks.to(self.device), metrics.to(self.device).float()
        so = self.student(x, metrics, masks, epoch, clusters=None)
        to = self.teacher(x, metrics, None, None, clusters=so["zero"], index=so["index"])
        grid = F.affine_grid(metrics[:, :2, :], size=(masks.shape[0], 1, masks.shape[1], masks.shape[2]), align_corners=False)
        warped = F.grid_sample(masks.unsqueeze(1), grid.to(masks.device), align_corners=False)
        so["gt"] = [masks, (warped > 0.1).float().squeeze()]
        return self.loss(so, to, epoch), so, to

    def step(self, x, masks, metrics, epoch=0):
        """train.py:229-272 (fp32, no GradScaler)."""
        u = self.ref.utils
        loss, so, to = self.forward(x, masks, metrics, epoch)
        self.opt.zero_grad()
        loss.backward()
        if self.clip_grad:
            u.clip_gradients(self.student, self.clip_grad)
        u.cancel_gradients_last_layer(epoch, self.student, self.freeze_last_layer)
        self.opt.step()
        with torch.no_grad():
            for mod in ("backbone", "head"):
                for q, k in zip(getattr(self.student, mod).parameters(), getattr(self.teacher, mod).parameters()):
                    k.data.mul_(self.m).add_((1 - self.m) * q.detach().data)
        return loss
