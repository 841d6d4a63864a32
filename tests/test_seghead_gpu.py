"""SegHead on the sm_100a implicit-GEMM path (ccd_conv_gemm + BatchNorm kernels) against PyTorch convolutions / the
oracle's seg_head_forward in fp32 (checkers only).  Reference: Dino/modules/segmentor.py:37-95."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from ccd_b200 import ops as o
    return o


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _nhwc(t):          # [N,C,H,W] -> [N*H*W, C]
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()


@pytest.mark.parametrize("C,H,W,n", [(384, 8, 32, 3), (192, 8, 32, 2), (128, 32, 128, 1)])
def test_conv3x3_fwd_dgrad_wgrad(ops, C, H, W, n):
    from ccd_b200.segmentor import CONV3_FWD_TAPS, CONV3_DGRAD_TAPS
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(n, C, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(128, C, 3, 3, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    M = n * H * W
    xm = _nhwc(x)
    wf = w.permute(0, 2, 3, 1).reshape(128, 9 * C).contiguous()
    out = torch.empty(M, 128, dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm(xm, wf, M, 128, 9 * C, ops.EPI_BF16, None, out, 128, 1, 1, H, W, C, 1, n, CONV3_FWD_TAPS, C)
    ref = F.conv2d(x.float(), w.float(), padding=1)
    assert _rel(out.float(), _nhwc(ref)) < 6e-3
    # data gradient
    dy = torch.randn(n, 128, H, W, device="cuda", generator=g).to(torch.bfloat16)
    wd = w.permute(1, 2, 3, 0).reshape(C, 9 * 128).contiguous()
    dx = torch.empty(M, C, dtype=torch.float32, device="cuda")
    ops.conv_gemm(_nhwc(dy), wd, M, C, 9 * 128, ops.EPI_F32, None, dx, C, 1, 1, H, W, 128, 1, n, CONV3_DGRAD_TAPS, 128)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    (F.conv2d(xr, wr, padding=1) * dy.float()).sum().backward()
    assert _rel(dx, _nhwc(xr.grad)) < 1e-3
    # weight gradient (split-K over positions)
    gw = torch.zeros(128, 9 * C, dtype=torch.float32, device="cuda")
    ops.conv_gemm(xm, _nhwc(dy), 128, 9 * C, M, ops.EPI_F32, None, gw, 9 * C, 4, 2, H, W, C, 1, n, CONV3_FWD_TAPS, C)
    assert _rel(gw.reshape(128, 3, 3, C).permute(0, 3, 1, 2), wr.grad) < 1e-3


@pytest.mark.parametrize("Cin,H,W,n", [(192, 8, 32, 2), (128, 16, 64, 2)])
def test_conv_transpose_fwd_dgrad_wgrad(ops, Cin, H, W, n):
    from ccd_b200.segmentor import convt_fwd_taps, convt_bwd_taps
    g = torch.Generator(device="cuda").manual_seed(Cin)
    x = torch.randn(n, Cin, H, W, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cin, 128, 4, 4, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(128, device="cuda", generator=g)
    M = n * H * W
    xm = _nhwc(x)
    out = torch.full((4 * M, 128), float("nan"), dtype=torch.bfloat16, device="cuda")
    for py in range(2):
        for px in range(2):
            taps, kk = convt_fwd_taps(py, px)
            wpar = torch.stack([w[:, :, ky, kx] for ky, kx in kk], 0).permute(2, 0, 1).reshape(128, 4 * Cin).contiguous()
            ops.conv_gemm(xm, wpar, M, 128, 4 * Cin, ops.EPI_BF16, b, out, 128, 1, 1, H, W, Cin, 1, n, taps, Cin, rowmap=1, py=py, px=px)
    ref = F.conv_transpose2d(x.float(), w.float(), b, stride=2, padding=1)
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), _nhwc(ref)) < 6e-3
    dy = torch.randn(n, 128, 2 * H, 2 * W, device="cuda", generator=g).to(torch.bfloat16)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    (F.conv_transpose2d(xr, wr, b, stride=2, padding=1) * dy.float()).sum().backward()
    tb = convt_bwd_taps(128)
    dym = _nhwc(dy)
    wd = w.permute(0, 2, 3, 1).reshape(Cin, 16 * 128).contiguous()
    dx = torch.empty(M, Cin, dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm(dym, wd, M, Cin, 16 * 128, ops.EPI_BF16, None, dx, Cin, 1, 1, H, W, 256, 2, n, tb, 128)
    assert _rel(dx.float(), _nhwc(xr.grad)) < 6e-3
    gw = torch.zeros(Cin, 16 * 128, dtype=torch.float32, device="cuda")
    ops.conv_gemm(dym, xm, Cin, 16 * 128, M, ops.EPI_F32, None, gw, 16 * 128, 3, 2, H, W, 256, 2, n, tb, 128)
    assert _rel(gw.reshape(Cin, 4, 4, 128).permute(0, 3, 1, 2), wr.grad) < 1e-3


def test_batchnorm_relu_fwd_bwd(ops):
    g = torch.Generator(device="cuda").manual_seed(7)
    M, C = 5000, 128
    z = (torch.randn(M, C, device="cuda", generator=g) * 2 + 0.5).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    sums = ops.bn_stats(z, C, M, C)
    mean, rstd = ops.bn_finalize(sums, M, 1e-5, 0.1, rm, rv, C)
    y = torch.empty(M, 192, dtype=torch.bfloat16, device="cuda")
    ops.bn_apply_relu(z, C, mean, rstd, gamma, beta, y[:, 64:], 192, M, C)
    zr = z.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, dtype=torch.double, device="cuda"), torch.ones(C, dtype=torch.double, device="cuda")
    ref = torch.relu(F.batch_norm(zr, rm2, rv2, gr, br, training=True, momentum=0.1, eps=1e-5))
    assert (y[:, 64:].double() - ref).abs().max() < 3e-2
    assert (rm.double() - rm2).abs().max() < 1e-4 and (rv.double() - rv2).abs().max() < 1e-3
    dy = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    (ref * dy.double()).sum().backward()
    s2 = ops.bn_bwd_reduce(dy, C, z, C, mean, rstd, gamma, beta, M, C)
    dz = ops.bn_bwd_apply(dy, C, z, C, mean, rstd, gamma, beta, s2, 1.0 / M, M, C)
    assert _rel(s2[:C], br.grad) < 2e-3 and _rel(s2[C:], gr.grad) < 2e-3
    assert _rel(dz.float(), zr.grad) < 1e-2


@pytest.mark.parametrize("variant", [1, 0], ids=["mma", "cuda-core"])
@pytest.mark.parametrize("n", [1, 3, 40])
def test_cls_conv_kernels(ops, n, variant):
    """cls = Conv2d(128 -> 2, 3x3, pad 1) (segmentor.py:88,94): forward, data and weight/bias gradients against F.conv2d
    in fp64, for the warp-MMA kernels (fp32 operands split into bf16 hi + lo) and the CUDA-core kernels."""
    ops.set_seg_cls_variant(variant)
    try:
        _cls_conv_case(ops, n)
    finally:
        ops.set_seg_cls_variant(1)


def _cls_conv_case(ops, n):
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(n, 128, 32, 128, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(2, 128, 3, 3, device="cuda", generator=g) * 0.05
    b = torch.randn(2, device="cuda", generator=g)
    xm = _nhwc(x)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=1)          # fp64: cuDNN's fp32 convolution may run in TF32
    out = ops.seg_cls_fwd(xm, w, b, n)
    assert out.shape == ref.shape and _rel(out, ref) < 1e-5
    dl = torch.randn(n, 2, 32, 128, device="cuda", generator=g)
    (ref * dl.double()).sum().backward()
    dx = ops.seg_cls_dgrad(dl, w, n)
    assert _rel(dx.float(), _nhwc(xr.grad)) < 4e-3          # bf16 output rounding
    dw, db = ops.seg_cls_wgrad(xm, dl, n)
    assert _rel(dw, wr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("E", [192, 384])
def test_seghead_against_oracle(E):
    """Whole SegHead forward + backward (training-mode BN) against the oracle restatement in fp32 on the same weights."""
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    from ccd_b200.segmentor import SegHead
    n = 4
    head = SegHead(in_channels=E).cuda()
    sd = S.fill_state_dict({k: v.shape for k, v in head.state_dict().items()}, 9, 0.05)
    head.load_state_dict(sd)
    head.train()
    g = torch.Generator(device="cuda").manual_seed(E)
    taps_store = [torch.randn(n * 256, E, device="cuda", generator=g).requires_grad_(True) for _ in range(3)]
    taps = [t.view(n, 8, 32, E).permute(0, 3, 1, 2) for t in taps_store]
    out = head(taps)
    dl = torch.randn(out.shape, device="cuda", generator=g)
    (out * dl).sum().backward()
    osd = {"segmentation." + k: v.cuda().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    rt = [t.detach().view(n, 8, 32, E).permute(0, 3, 1, 2).clone().requires_grad_(True) for t in taps_store]
    stats = {}
    ref = O.seg_head_forward(osd, "segmentation.", rt, stats)
    (ref * dl).sum().backward()
    assert out.shape == ref.shape == (n, 2, 32, 128)
    assert (out - ref).abs().max() < 0.05 * ref.abs().max() + 1e-3
    assert _rel(out, ref) < 2e-2
    # bf16 operands through 5 conv + 8 BatchNorm/ReLU layers (mask flips at zero crossings): ~10 % relative gradient
    # noise, the level stock bf16-autocast cuDNN shows on the same head (tests/grad_parity_report.py: cos 0.985-0.99)
    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return (a @ b / (a.norm() * b.norm())).item()
    for a, b in zip(taps_store, rt):
        assert cos(a.grad, b.grad.permute(0, 2, 3, 1).reshape(-1, E)) > 0.985
    bad = []
    for k, p in head.named_parameters():
        if k.startswith("conv_mla"):
            assert p.grad is None
            continue
        r = osd["segmentation." + k].grad
        if r.norm() < 1e-6 or k in ("unpool1.0.bias", "unpool2.0.bias"):
            continue        # a bias in front of a training-mode BatchNorm has zero gradient: both sides are rounding noise
        if cos(p.grad, r) < 0.985:
            bad.append((k, cos(p.grad, r)))
    assert not bad, bad
    # running statistics follow nn.BatchNorm2d(momentum=0.1)
    mu, var = stats["segmentation.unpool2.1"]
    assert (head.unpool2[1].running_mean - (0.9 * sd["unpool2.1.running_mean"].cuda() + 0.1 * mu)).abs().max() < 2e-2
    assert int(head.unpool2[1].num_batches_tracked) == 1
