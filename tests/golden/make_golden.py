"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) in the build container.

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these fixtures are what pins both the
oracle restatement (oracle/ccd_oracle.py) and, through it, the CUDA path.  Inputs and weights are regenerated
from seeds by ccd_b200.synthetic (name-keyed deterministic weights), so only OUTPUTS are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_import  # noqa: E402
from ccd_b200 import synthetic as S  # noqa: E402

sys.path.insert(0, HERE)
from cases import CASES, COL_STRIDE, COL_STRIDE_OF, EPOCH, GRAD_HEAD, SEG_KEEP  # noqa: E402


def compact_from_onehot(z):
    """[N,26,H,W] one-hots -> uint8 [N,H,W] slot+1 (0 = background).  Lossless: slots are disjoint."""
    z = z.cpu().numpy()
    idx = np.arange(1, 27, dtype=np.float32)[None, :, None, None]
    assert (z.sum(1) <= 1).all()
    return (z * idx).sum(1).astype(np.uint8)


def run_case(ref, name, arch, E, B, K, sseed, tseed, std, norm_last):
    vit = getattr(ref.vits, arch)
    student = ref.dv.ABIDINOModel(vit(patch_size=4, drop_path_rate=0.0), ref.seg.SegHead(in_channels=E),
                                  ref.vits.DINOHead(E, K, norm_last_layer=norm_last))
    teacher = ref.dv.ABIDINOModel(vit(patch_size=4), None, ref.vits.DINOHead(E, K))
    student.load_state_dict(S.fill_state_dict({k: v.shape for k, v in student.state_dict().items()}, sseed, std))
    teacher.load_state_dict(S.fill_state_dict({k: v.shape for k, v in teacher.state_dict().items()}, tseed, std))
    for p in teacher.parameters():
        p.requires_grad = False
    x, masks, metrics = S.make_batch(B, seed=1234)
    loss_mod = ref.loss.DINOLoss(K, 2, 0.04, 0.04, 0, 10 if EPOCH.get(name, 0) < 10 else 101)
    center0 = 0.01 * torch.randn(1, K, generator=torch.Generator().manual_seed(5))
    loss_mod.center.copy_(center0)

    epoch = EPOCH.get(name, 0)
    stride = COL_STRIDE_OF.get(name, COL_STRIDE)
    so = student(x, metrics, masks, epoch, clusters=None)                               # train.py:232
    to = teacher(x, metrics, None, None, clusters=so["zero"], index=so["index"])        # train.py:233
    ag = F.affine_grid(metrics[:, :2, :], size=(B, 1, 32, 128))                         # train.py:234-236
    mi = (F.grid_sample(masks.unsqueeze(1), ag) > 0.1).float().squeeze()
    so["gt"] = [masks, mi]
    loss = loss_mod(so, to, epoch)                                                      # train.py:238
    loss.backward()

    out = {
        "loss": loss.item(), "mask_loss": loss_mod.last_losses["mask_loss"].item(),
        "dino_loss": loss_mod.last_losses["Dino_loss"].item(),
        "student_logits": so["instances_view"].detach()[:, ::stride].numpy(),
        "teacher_logits": to["instances_view"].detach()[:, ::stride].numpy(),
        "seg_logits": (so["mask"].detach() if name not in SEG_KEEP else
                       torch.cat([so["mask"].detach()[:SEG_KEEP[name]], so["mask"].detach()[B:B + SEG_KEEP[name]]])
                       ).numpy().astype(np.float16 if epoch < 30 else np.float32),
        "clusters_compact": compact_from_onehot(so["zero"]),
        "new_index": so["index"].numpy(),
        "gt_warped": mi.numpy().astype(np.uint8),
        "center_after": loss_mod.center[:, ::stride].numpy(),
        "teacher_feature": to["feature"].detach()[:, ::7, :, ::3].numpy(),
    }
    names, norms, heads = [], [], []
    for n, p in student.named_parameters():
        if p.grad is None:
            continue
        names.append(n)
        norms.append(p.grad.norm().item())
        h = p.grad.flatten()[:GRAD_HEAD].numpy()
        heads.append(np.pad(h, (0, GRAD_HEAD - len(h))))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    out["grad_heads"] = np.stack(heads)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"], "rows", so["instances_view"].shape[0], "grads", len(names))


def run_ccl(ref):
    """label_cluster (Dino/utils/DBSCAN.py:61-103) on irregular masks, incl. empty / full / >26 components."""
    masks = torch.cat([S.random_masks(40, seed=11, blobs=True), S.random_masks(8, seed=12, blobs=False, density=0.55),
                       S.random_masks(8, seed=13, blobs=False, density=0.2), torch.zeros(1, 32, 128),
                       torch.ones(1, 32, 128), S.make_batch(6)[1]])
    # many small components: > 26 candidates with area >= 30 is impossible in 4096 px only if small; build stripes
    stripes = torch.zeros(2, 32, 128)
    for c in range(40):
        stripes[0, 0:16, 3 * c: 3 * c + 2] = 1.0            # 40 components of 32 px (>26 survive the area rule)
        stripes[1, (c % 2) * 16:(c % 2) * 16 + 15, 3 * c: 3 * c + 2] = 1.0
    masks = torch.cat([masks, stripes])
    label = ref.dbscan.label_cluster()
    comp = np.zeros((masks.shape[0], 32, 128), dtype=np.uint8)
    for i, m in enumerate(masks.numpy()):
        z = label(m).astype(np.float32)
        comp[i] = compact_from_onehot(torch.tensor(z[None]))[0]
    np.savez_compressed(os.path.join(HERE, "ccl_cases.npz"), masks=masks.numpy().astype(np.uint8), compact=comp)
    print("ccl cases", masks.shape[0], "max slots", comp.max())


def main():
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference()
    ref_import.ensure_gloo_group()
    torch.manual_seed(0)
    only = sys.argv[1:]
    if not only:
        run_ccl(ref)
    for name, cfg in CASES.items():
        if not only or name in only:
            run_case(ref, name, *cfg)


if __name__ == "__main__":
    main()
