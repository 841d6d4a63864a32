"""Two ranks of the PRODUCT step (train.py:93-110 wrapping: SyncBN conversion + DDP; each rank holds its own ragged half of the
batch) against the oracle on the concatenated batch.  What is pinned:
  * the DDP-averaged student gradients  ==  d/dtheta [ (L_0 + L_1) / 2 ]  with the SegHead's BatchNorm statistics taken over
    BOTH ranks' rows (SyncBatchNorm) and every per-rank mean (pixels, character rows R_r) taken per rank, as the reference does;
  * DINOLoss.center after update_center on each rank: SUM over ranks of the teacher row sums divided by local_rows * world
    (Dino/loss/Dino_loss.py:133-143) -- ranks with different R_r end up with DIFFERENT centres, replicated here.
Both ranks share cuda:0 and talk over gloo, so the test runs on a 1-GPU box (same arrangement as test_syncbn_2rank_gpu.py)."""
import os
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ARCH, E, K, B = "vit_tiny", 192, 4096, 4          # per-rank batch 4; rank 1 takes samples 4..7 (more characters -> R_1 > R_0)


def _setup():
    sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE]
    from ccd_b200 import synthetic as S
    from test_pretrain_parity_gpu import build
    student, teacher, ssd, tsd = build(ARCH, E, K, 31, 32, 0.05, False)
    x, masks, metrics = S.make_batch(2 * B, seed=4321)
    center0 = 0.01 * torch.randn(1, K, generator=torch.Generator().manual_seed(5))
    return student, teacher, ssd, tsd, x, masks, metrics, center0


def _worker(rank, world, init_file, out_dir):
    import torch.distributed as dist
    import torch.nn as nn
    student, teacher, ssd, tsd, x, masks, metrics, center0 = _setup()
    from Dino.loss.Dino_loss import DINOLoss
    from ccd_b200 import ops
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    student = nn.SyncBatchNorm.convert_sync_batchnorm(student)                                       # train.py:96-98
    student = nn.parallel.DistributedDataParallel(student, device_ids=[0], find_unused_parameters=True)   # train.py:106
    sl = slice(rank * B, (rank + 1) * B)
    xr, mr, tr = x[sl].cuda(), masks[sl].cuda(), metrics[sl].cuda()
    crit = DINOLoss(K, 2, 0.04, 0.04, 0, 101).cuda()
    crit.center.copy_(center0)
    so = student(xr, tr, mr, 0, clusters=None)
    to = teacher(xr, tr, None, None, clusters=so["zero"], index=so["index"])
    so["gt"] = [mr, ops.warp_mask(mr, tr)]
    loss = crit(so, to, 0)
    loss.backward()
    torch.cuda.synchronize()
    torch.save({"loss": loss.item(), "rows": so["instances_view"].shape[0], "center": crit.center.cpu(),
                "grads": {n: p.grad.cpu() for n, p in student.module.named_parameters() if p.grad is not None}},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_product_step_equals_oracle_on_concatenated_batch():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        r = [torch.load(os.path.join(d, f"rank{i}.pt")) for i in range(2)]
    _, _, ssd, tsd, x, masks, metrics, center0 = _setup()
    import ccd_oracle as O
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in ssd.items()}
    # encoder per rank (no cross-sample coupling), SegHead over BOTH ranks' rows (SyncBN), everything else per rank
    toks, taps = [], []
    for k in range(2):
        sl = slice(k * B, (k + 1) * B)
        t, tp = O.vit_forward(sd, "backbone.", torch.cat([x[sl, 1], x[sl, 2]]), ARCH)
        toks.append(t); taps.append(tp)
    seg_all = O.seg_head_forward(sd, "segmentation.", [torch.cat([taps[0][i], taps[1][i]]) for i in range(3)])
    losses, zts, rows = [], [], []
    for k in range(2):
        sl = slice(k * B, (k + 1) * B)
        seg = seg_all[k * 2 * B:(k + 1) * 2 * B]
        src = torch.tensor(__import__("numpy").stack([O.label_cluster(m)[0] for m in masks[sl].numpy()]), dtype=torch.float32)
        clusters = torch.cat([src, O.warp_affine_binary(src, metrics[sl])])
        pooled, index = O.char_pool(toks[k], clusters)
        rws, _ = O.ragged_select(pooled, index)
        zs = O.dino_head_forward(sd, "head.", rws)
        with torch.no_grad():
            tt, _ = O.vit_forward(tsd, "backbone.", torch.cat([x[sl, 1], x[sl, 2]]), ARCH)
            pt, it = O.char_pool(tt, clusters)
            zt = O.dino_head_forward(tsd, "head.", O.ragged_select(pt, it)[0])
        gt = torch.cat([masks[sl], O.warp_affine_binary(masks[sl].unsqueeze(1), metrics[sl]).squeeze(1)])
        losses.append(O.seg_loss(seg, gt) + O.dino_distill_loss(zs, zt, center0, 0.04))
        zts.append(zt); rows.append(zs.shape[0])
    (0.5 * (losses[0] + losses[1])).backward()
    assert rows[0] != rows[1] and [r[0]["rows"], r[1]["rows"]] == rows                 # ragged, and the same raggedness
    for k in range(2):
        assert abs(r[k]["loss"] - losses[k].item()) / losses[k].item() <= 1e-3, (k, r[k]["loss"], losses[k].item())
    # F8: every rank divides the global SUM by ITS OWN row count
    world_sum = zts[0].sum(0, keepdim=True) + zts[1].sum(0, keepdim=True)
    for k in range(2):
        want = O.updated_center(center0, zts[k], world_sum=world_sum, world_size=2)
        assert (r[k]["center"] - want).abs().max() <= 3e-4, k
    assert (r[0]["center"] - r[1]["center"]).abs().max() > 1e-6                        # the reference quirk: centres differ
    # DDP leaves the SAME averaged gradient on both ranks; it equals the oracle's gradient of the mean of the two rank losses

    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return (a @ b / (a.norm() * b.norm() + 1e-30)).item()
    bad = []
    for n, g0 in r[0]["grads"].items():
        assert torch.equal(g0, r[1]["grads"][n]), n
        want = sd[n].grad
        if want is None or want.norm() < 1e-7:
            continue
        c = cos(g0, want)
        ratio = (g0.norm() / want.norm()).item()
        if c < (0.999 if n.startswith("head.") else 0.98) or not 0.9 < ratio < 1.1:
            bad.append((n, round(c, 4), round(ratio, 3)))
    assert len(bad) <= 3, bad                                                          # BN-nullified conv biases: ~1e-9 norms
