"""SegHead under SyncBatchNorm with two ranks (train.py:96-101): each rank holds half of the batch, the BatchNorm statistics
and their gradients are exchanged (one all-reduce per independent group of layers), and the result must equal the
single-process oracle on the whole batch.  Both ranks share cuda:0 and talk over gloo, so the test runs on a 1-GPU box."""
import os
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
E, N_IMG = 192, 4


def _inputs():
    from ccd_b200 import synthetic as S
    from ccd_b200.segmentor import SegHead
    head = SegHead(in_channels=E)
    sd = S.fill_state_dict({k: v.shape for k, v in head.state_dict().items()}, 11, 0.05)
    g = torch.Generator().manual_seed(3)
    taps = [torch.randn(N_IMG * 256, E, generator=g) for _ in range(3)]
    dl = torch.randn(N_IMG, 2, 32, 128, generator=g)
    return sd, taps, dl


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import torch.nn as nn
    from ccd_b200.segmentor import SegHead
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    sd, taps, dl = _inputs()
    head = SegHead(in_channels=E)
    head.load_state_dict(sd)
    head = nn.SyncBatchNorm.convert_sync_batchnorm(head).cuda().train()
    per = N_IMG // world
    rows = slice(rank * per * 256, (rank + 1) * per * 256)
    ts = [t[rows].cuda().requires_grad_(True) for t in taps]
    out = head([t.view(per, 8, 32, E).permute(0, 3, 1, 2) for t in ts])
    (out * dl[rank * per:(rank + 1) * per].cuda()).sum().backward()
    torch.save({"out": out.detach().cpu(), "dtaps": [t.grad.cpu() for t in ts],
                "grads": {k: p.grad.cpu() for k, p in head.named_parameters() if p.grad is not None},
                "rm": head.unpool2[1].running_mean.cpu()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_seghead_syncbn_two_ranks_equals_full_batch_oracle():
    import torch.multiprocessing as mp
    import ccd_oracle as O
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        r = [torch.load(os.path.join(d, f"rank{i}.pt")) for i in range(2)]
    sd, taps, dl = _inputs()
    osd = {"segmentation." + k: v.cuda().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    rt = [t.cuda().view(N_IMG, 8, 32, E).permute(0, 3, 1, 2).clone().requires_grad_(True) for t in taps]
    stats = {}
    ref = O.seg_head_forward(osd, "segmentation.", rt, stats)
    (ref * dl.cuda()).sum().backward()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return (a @ b / (a.norm() * b.norm())).item()
    out = torch.cat([r[0]["out"], r[1]["out"]]).cuda()
    assert rel(out, ref) < 2e-2
    for j in range(3):
        got = torch.cat([r[0]["dtaps"][j], r[1]["dtaps"][j]]).cuda()
        assert cos(got, rt[j].grad.permute(0, 2, 3, 1).reshape(-1, E)) > 0.985
    for k, g0 in r[0]["grads"].items():
        want = osd["segmentation." + k].grad
        if want.norm() < 1e-6 or k in ("unpool1.0.bias", "unpool2.0.bias"):
            continue
        assert cos((g0 + r[1]["grads"][k]).cuda(), want) > 0.985, k          # the loss is a sum over the global batch
    mu, _ = stats["segmentation.unpool2.1"]
    for i in range(2):                                                        # both ranks track the GLOBAL batch statistics
        assert (r[i]["rm"].cuda() - (0.9 * sd["unpool2.1.running_mean"].cuda() + 0.1 * mu)).abs().max() < 2e-2
