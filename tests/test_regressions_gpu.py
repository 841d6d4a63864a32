"""GPU regression tests for the round-1 review findings: optimizer checkpointing in mid-training, GEMM-weight copies
after `.data` updates (the reference's own EMA loop), the distillation-loss gradient for logits that do not come from this
repository's head, logits feeding two losses, SegHead backward in eval mode."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _tiny(K=1024, seed=1):
    from Dino.model.dino_vision import ABIDINOModel
    from Dino.modules import vision_transformer as vits
    from Dino.modules.segmentor import SegHead
    from ccd_b200 import synthetic as S
    m = ABIDINOModel(vits.vit_tiny(patch_size=4), SegHead(in_channels=192), vits.DINOHead(192, K, norm_last_layer=False))
    m.load_state_dict(S.fill_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed, 0.05))
    return m.cuda()


def test_adamw_state_dict_between_steps_keeps_stepping():
    """ADVICE high #1: train.py saves optimizer.state_dict() every pseudo-epoch (train.py:197-203); the next step must work and the
    trajectory must stay torch.optim.AdamW's."""
    from ccd_b200.optim import AdamW
    torch.manual_seed(0)
    ps = [nn.Parameter(torch.randn(300, 70, device="cuda")), nn.Parameter(torch.randn(513, device="cuda"))]
    qs = [nn.Parameter(p.detach().clone()) for p in ps]
    ours, ref = AdamW(ps, lr=1e-2, weight_decay=0.1), torch.optim.AdamW(qs, lr=1e-2, weight_decay=0.1)
    for it in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        ours.step(); ref.step()
        sd = ours.state_dict()                                    # must not touch the live state
        assert all(isinstance(st["step"], int) for st in ours.state.values())
        assert all(torch.is_tensor(st["step"]) for st in sd["state"].values())
    for p, q in zip(ps, qs):
        assert (p - q).abs().max() <= 1e-5
    # and the saved state loads into torch's optimizer and back
    ref2 = torch.optim.AdamW([nn.Parameter(q.detach().clone()) for q in qs], lr=1e-2, weight_decay=0.1)
    ref2.load_state_dict(sd)
    ours.load_state_dict(ref.state_dict())
    for p in ps:
        p.grad = torch.ones_like(p)
    ours.step()


def test_data_ema_updates_reach_the_gemm_operands():
    """ADVICE high #2: the reference's EMA is `param_k.data.mul_(m).add_(...)` (train.py:264-272), invisible to tensor version
    counters.  After it, the teacher's forward must use the NEW weights in its GEMMs (bf16 operand copies re-cast)."""
    from ccd_b200 import synthetic as S
    student, teacher = _tiny(seed=1), _tiny(seed=2)
    x, masks, metrics = S.make_batch(4, seed=3, device="cuda")
    with torch.no_grad():
        cm = student(x, metrics, masks, 0)["zero"]
        before = teacher(x, metrics, None, None, clusters=cm)["instances_view"].clone()
        for mod in ("backbone", "head"):
            for q, k in zip(getattr(student, mod).parameters(), getattr(teacher, mod).parameters()):
                k.data.mul_(0.3).add_((1 - 0.3) * q.detach().data)
        after = teacher(x, metrics, None, None, clusters=cm)["instances_view"]
        fresh = _tiny(seed=2)
        fresh.load_state_dict(teacher.state_dict())
        want = fresh(x, metrics, None, None, clusters=cm)["instances_view"]
    assert (after - before).abs().max() > 1e-2                     # the update is visible ...
    assert torch.equal(after, want)                                # ... and identical to a model built from the updated weights


def test_dino_loss_gradient_for_foreign_logits_and_two_losses():
    """ADVICE medium: logits that do not come from ccd_b200's DINOHead get an ordinary dense gradient; logits that also feed a
    second loss keep BOTH gradients through the head."""
    import ccd_oracle as O
    from Dino.loss.Dino_loss import DINOLoss
    from ccd_b200.loss import DinoCEFn
    K, R2 = 2048, 12
    g = torch.Generator().manual_seed(0)
    zs0, zt, c = torch.randn(R2, K, generator=g), torch.randn(R2, K, generator=g), 0.01 * torch.randn(1, K, generator=g)
    # (a) a leaf tensor
    zs = zs0.clone().cuda().requires_grad_(True)
    L = DinoCEFn.apply(zs, zt.cuda(), c.cuda(), 0.1, 0.04)
    L.backward()
    ref = zs0.clone().requires_grad_(True)
    O.dino_distill_loss(ref, zt, c, 0.04).backward()
    assert zs.grad is not None and (zs.grad.cpu() - ref.grad).abs().max() <= 2e-2 * ref.grad.abs().max() + 1e-7
    assert zs.grad.abs().sum() > 0
    # (b) a scaled view of head logits + (c) head logits feeding two losses
    from ccd_b200.head import DINOHead, DLOGITS_STASH, HEAD_LOGITS
    torch.manual_seed(1)
    head = DINOHead(192, K, norm_last_layer=False).cuda()
    rows = torch.randn(R2, 192, device="cuda")

    def grads(fn):
        head.zero_grad()
        fn(head(rows)).backward()
        return {n: p.grad.clone() for n, p in head.named_parameters() if p.grad is not None}

    ztc, cc = zt.cuda(), c.cuda()
    g_ce = grads(lambda z: DinoCEFn.apply(z, ztc, cc, 0.1, 0.04))
    g_scaled = grads(lambda z: DinoCEFn.apply(z * 1.0, ztc, cc, 0.1, 0.04))                 # z * 1.0 is NOT the registered tensor
    g_l2 = grads(lambda z: 1e-3 * (z ** 2).mean())
    g_both = grads(lambda z: DinoCEFn.apply(z, ztc, cc, 0.1, 0.04) + 1e-3 * (z ** 2).mean())
    for n in g_ce:
        ref_n = g_ce[n].norm() + 1e-12
        assert (g_scaled[n] - g_ce[n]).norm() <= 2e-2 * ref_n, n
        assert (g_both[n] - (g_ce[n] + g_l2[n])).norm() <= 2e-2 * (g_ce[n] + g_l2[n]).norm() + 1e-9, n
    assert not DLOGITS_STASH and not HEAD_LOGITS                                            # nothing left behind


def test_seghead_eval_mode_backward_uses_running_statistics():
    """ADVICE low: with the head in eval mode the BN layers are affine maps of the running statistics; the backward must not apply
    the batch-statistic correction terms."""
    from ccd_b200 import synthetic as S
    from ccd_b200.segmentor import SegHead
    E, n = 192, 3
    head = SegHead(in_channels=E)
    head.load_state_dict(S.fill_state_dict({k: v.shape for k, v in head.state_dict().items()}, 5, 0.05))
    head = head.cuda().eval()
    ref = copy.deepcopy(head).float()
    g = torch.Generator().manual_seed(2)
    taps = [torch.randn(n, E, 8, 32, generator=g).cuda() for _ in range(3)]
    dl = torch.randn(n, 2, 32, 128, generator=g).cuda()
    a = [t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True) for t in taps]     # NHWC storage, NCHW view
    out = head(a)
    (out * dl).sum().backward()
    b = [t.clone().requires_grad_(True) for t in taps]
    torch.backends.cudnn.allow_tf32 = False
    h = ref.mlahead
    y = torch.cat([br(t) for br, t in zip((h.head2, h.head3, h.head4), b)], 1)
    want = ref.cls(ref.unpool2(ref.unpool1(y)))
    (want * dl).sum().backward()

    def cos(u, v):
        u, v = u.double().flatten(), v.double().flatten()
        return (u @ v / (u.norm() * v.norm() + 1e-30)).item()
    assert ((out - want).norm() / want.norm()).item() < 2e-2
    # same bar as the training-mode SegHead tests (bf16 activations through 8 layers, test_syncbn_2rank_gpu.py); applying the
    # training-mode BatchNorm backward here instead drops these cosines to ~0.9 and below
    cs = [cos(u.grad, v.grad) for u, v in zip(a, b)]
    assert min(cs) > 0.985, cs
    gp, gr = dict(head.named_parameters()), dict(ref.named_parameters())
    for k in ("mlahead.head2.0.weight", "mlahead.head3.1.weight", "unpool1.0.weight", "unpool2.1.bias", "cls.weight"):
        assert cos(gp[k].grad, gr[k].grad) > 0.985, (k, cos(gp[k].grad, gr[k].grad))
    # the training-mode formula on the same inputs is measurably different (so the test can tell the two apart)
    head.train()
    c = [t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True) for t in taps]
    (head(c) * dl).sum().backward()
    assert min(cos(u.grad, v.grad) for u, v in zip(c, b)) < min(cs) - 0.01
