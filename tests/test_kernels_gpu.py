"""Per-kernel parity tests on the GPU, every call going through the C ABI (ccd_b200.ops -> libccd_b200.so).
The checker is the oracle restatement (oracle/ccd_oracle.py) or the same contraction in fp32/fp64 torch on the
bf16-rounded operands.  Tolerances are stated per test; integer / index kernels are bit-exact."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ops():
    from ccd_b200 import ops as o
    return o


def _bf(t):
    return t.to(torch.bfloat16)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


# ---------------------------------------------------------------------------------------------------------
# GEMM: every majorness combination and epilogue
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(params=[0, 2, 1], ids=["smem-transpose", "thread-per-row", "auto"])
def gemm_epilogue(ops, request):
    """Both full-tile epilogue implementations of the persistent GEMM (and the per-shape default)."""
    ops.set_gemm_epilogue(request.param)
    yield request.param
    ops.set_gemm_epilogue(1)


@pytest.mark.parametrize("M,N,K", [(256, 384, 384), (512, 1152, 192), (52, 256, 2048), (300, 192, 448), (128, 128, 64)])
def test_gemm_kmajor_bf16_bias(ops, M, N, K, gemm_epilogue):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = _bf(torch.randn(M, K, device="cuda", generator=g))
    B = _bf(torch.randn(N, K, device="cuda", generator=g) * 0.1)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_BF16, bias, out)
    ref = A.float() @ B.float().t() + bias
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), ref) < 4e-3        # bf16 output rounding


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(384, 1536, 1024), (192, 64, 520), (256, 256, 52)])
def test_gemm_majorness_f32(ops, a_mn, b_mn, M, N, K, gemm_epilogue):
    g = torch.Generator(device="cuda").manual_seed(11 + a_mn * 2 + b_mn)
    A = _bf(torch.randn(M, K, device="cuda", generator=g))
    B = _bf(torch.randn(N, K, device="cuda", generator=g))
    Amem = A.t().contiguous() if a_mn else A
    Bmem = B.t().contiguous() if b_mn else B
    if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
        pytest.skip("leading dimension must be a multiple of 8 elements")
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    ops.gemm(Amem, Bmem, M, N, K, a_mn, b_mn, ops.EPI_F32, None, out)
    ref = A.double() @ B.double().t()
    assert _rel(out, ref) < 1e-5


def test_gemm_splitk_atomic(ops):
    M, N, K = 384, 192, 8192
    g = torch.Generator(device="cuda").manual_seed(3)
    A = _bf(torch.randn(K, M, device="cuda", generator=g))      # MN-major storage [K, M]
    B = _bf(torch.randn(K, N, device="cuda", generator=g))
    out = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A, B, M, N, K, 1, 1, ops.EPI_F32, None, out, splits=16)
    ref = A.double().t() @ B.double()
    assert _rel(out, ref) < 1e-5


@pytest.mark.parametrize("N", [768, 1152, 200])      # N-tile 256 / 192 / 128 (ragged, column masking)
def test_gemm_epilogues(ops, N, gemm_epilogue):
    import ccd_oracle as O
    M, K = 512, 384
    g = torch.Generator(device="cuda").manual_seed(5)
    A = _bf(torch.randn(M, K, device="cuda", generator=g))
    B = _bf(torch.randn(N, K, device="cuda", generator=g) * 0.05)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    acc = A.float() @ B.float().t() + bias
    # GELU: pre-activation + activation
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    act = torch.empty_like(pre)
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_GELU, bias, pre, act)
    assert _rel(pre.float(), acc) < 4e-3
    assert (act.float() - O.gelu(pre.float())).abs().max() < 2e-2
    assert _rel(act.float(), O.gelu(pre.float())) < 4e-3
    act2 = torch.empty_like(pre)                      # inference form: no pre-activation output
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_GELU, bias, None, act2)
    assert torch.equal(act2, act)
    # residual
    res = torch.randn(M, N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_RESID, bias, out, None, res)
    assert _rel(out, acc + res) < 1e-5
    scale = torch.rand(M // 256, device="cuda", generator=g)          # DropPath scale per 256-row sequence
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_RESID, bias, out, None, res, seq_scale=scale)
    assert _rel(out, acc * scale.repeat_interleave(256)[:, None] + res) < 1e-5
    # dGELU
    hpre = _bf(torch.randn(M, N, device="cuda", generator=g))
    out_b = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    csum = torch.zeros(N, device="cuda")
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_DGELU, None, out_b, csum, hpre)          # out1 = fused column sums
    x = hpre.float().requires_grad_(True)
    O.gelu(x).sum().backward()
    assert _rel(out_b.float(), (acc - bias) * x.grad) < 5e-3
    assert _rel(csum, ((acc - bias) * x.grad).sum(0)) < 2e-3
    # positional epilogue
    pos = torch.randn(256, N, device="cuda", generator=g)
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_POS, bias, out, None, pos)
    assert _rel(out, acc + pos.repeat(M // 256, 1)) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------
def _attn_ref(qkv, S, H):
    E = H * 64
    q, k, v = qkv.double().reshape(S, 256, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * 0.125
    p = torch.softmax(s, -1)
    o = (p @ v).transpose(1, 2).reshape(S * 256, E)
    lse2 = torch.logsumexp(s, -1) * math.log2(math.e)
    return o, lse2, p


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("S,H", [(2, 3), (5, 6), (70, 6)])
def test_mhsa_fwd(ops, variant, S, H):
    g = torch.Generator(device="cuda").manual_seed(S * 10 + H)
    qkv = _bf(torch.randn(S * 256, 3 * H * 64, device="cuda", generator=g) * 1.5)
    o, lse = ops.mhsa_fwd(qkv, S, H, True, variant)
    ro, rl, _ = _attn_ref(qkv, S, H)
    assert torch.isfinite(o.float()).all()
    assert (o.double() - ro).abs().max() < 2e-2          # bf16 P and bf16 output
    assert _rel(o.float(), ro) < 1e-2
    assert (lse.double() - rl).abs().max() < 2e-2


@pytest.mark.parametrize("variant", [1, 0])           # 1 = pipelined persistent kernel (default), 0 = first version
@pytest.mark.parametrize("S,H", [(2, 3), (3, 8), (70, 6)])   # 420 items > 148 SMs: several items per persistent CTA
def test_mhsa_bwd(ops, S, H, variant):
    g = torch.Generator(device="cuda").manual_seed(S + H)
    qkv = _bf(torch.randn(S * 256, 3 * H * 64, device="cuda", generator=g))
    d_o = _bf(torch.randn(S * 256, H * 64, device="cuda", generator=g))
    o, lse = ops.mhsa_fwd(qkv, S, H, True, 0)
    dbias = torch.zeros(3 * H * 64, device="cuda")
    ops.set_mhsa_bwd_variant(variant)
    try:
        dqkv = ops.mhsa_bwd(qkv, o, d_o, lse, S, H, dbias=dbias)
        torch.cuda.synchronize()
    finally:
        ops.set_mhsa_bwd_variant(1)
    x = qkv.double().requires_grad_(True)
    ro, _, _ = _attn_ref(x, S, H)
    (ro * d_o.double()).sum().backward()
    assert torch.isfinite(dqkv.float()).all()
    E = H * 64
    for name, sl in (("dq", slice(0, E)), ("dk", slice(E, 2 * E)), ("dv", slice(2 * E, 3 * E))):
        r = _rel(dqkv[:, sl].float(), x.grad[:, sl])
        assert r < 2e-2, (name, r)
    # qkv-bias gradient = column sums of dqkv (fused: q from the dQ tiles, v = column sums of d_o, k identically zero)
    want = x.grad.sum(0)
    scale = want.abs().max().item()
    assert (dbias.double() - want).abs().max().item() < 2e-2 * scale, ((dbias.double() - want).abs().max().item(), scale)
    assert want[E:2 * E].abs().max().item() < 1e-6 * scale          # the key part vanishes analytically


def test_mhsa_bwd_bias_gradient_from_proj_bias_gradient(ops):
    """The v part of the qkv-bias gradient as the encoder obtains it: d_o = dY W_proj, so colsum(d_o) = colsum(dY) W_proj
    (ccd_vecmat_add_f32 inside ccd_mhsa_bwd) -- against autograd of the attention with a v-bias."""
    S, H = 3, 6
    E = H * 64
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = _bf(torch.randn(S * 256, 3 * E, device="cuda", generator=g))
    dY = torch.randn(S * 256, E, device="cuda", generator=g)
    W = torch.randn(E, E, device="cuda", generator=g) * 0.05
    d_o = _bf(dY @ W)
    o, lse = ops.mhsa_fwd(qkv, S, H, True, 0)
    dbias = torch.zeros(3 * E, device="cuda")
    dqkv = ops.mhsa_bwd(qkv, o, d_o, lse, S, H, dbias=dbias, dproj_bias=dY.sum(0).contiguous(), w_proj=W.contiguous())
    torch.cuda.synchronize()
    want = dqkv.double().sum(0)
    scale = want.abs().max().item()
    assert (dbias.double() - want).abs().max().item() < 2e-2 * scale
    out = torch.zeros(E, device="cuda")
    ops.vecmat_add(dY.sum(0).contiguous(), W.contiguous(), out)
    assert _rel(out, dY.double().sum(0) @ W.double()) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# row-wise kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E", [192, 384, 512])
def test_layernorm_fwd_bwd(ops, E):
    import ccd_oracle as O
    rows = 1000
    g = torch.Generator(device="cuda").manual_seed(E)
    x = torch.randn(rows, E, device="cuda", generator=g) * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(E, device="cuda", generator=g)
    beta = 0.1 * torch.randn(E, device="cuda", generator=g)
    yb, yf = ops.layernorm_fwd(x, gamma, beta, True, True)
    xr = x.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    ref = O.layer_norm(xr, gr, br)
    assert (yf.double() - ref).abs().max() < 1e-4
    assert (yb.double() - ref).abs().max() < 3e-2
    dy = torch.randn(rows, E, device="cuda", generator=g)
    resid = torch.randn(rows, E, device="cuda", generator=g)
    (ref * dy.double()).sum().backward()
    for dyt in (dy, _bf(dy)):
        dgam = torch.zeros(E, device="cuda")
        dbet = torch.zeros(E, device="cuda")
        dnext = torch.zeros(E, device="cuda")
        seq_scale = 0.5 + torch.rand((rows + 255) // 256, device="cuda", generator=g)
        dxf, dxb = ops.layernorm_bwd(x, gamma, dyt.contiguous(), resid, dgam, dbet, bf16_seq_scale=seq_scale, dbias_next=dnext)
        assert _rel(dnext, dxb.double().sum(0)) < 1e-5              # fused bias-gradient column sums of the bf16 copy
        dxb = (dxb.float() / seq_scale.repeat_interleave(256)[:rows, None]).to(torch.bfloat16)
        tol = 1e-4 if dyt.dtype == torch.float32 else 1e-2
        assert _rel(dxf - resid, xr.grad) < tol
        assert _rel(dgam, gr.grad) < tol and _rel(dbet, br.grad) < tol
        assert _rel(dxb.float(), dxf) < 8e-3


def test_colsum_and_norms(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _bf(torch.randn(3000, 1152, device="cuda", generator=g))
    out = torch.zeros(1152, device="cuda")
    ops.colsum_bf16(x, out)
    assert _rel(out, x.double().sum(0)) < 1e-5
    xf = torch.randn(777, 4096, device="cuda", generator=g)
    out = torch.zeros(4096, device="cuda")
    ops.colsum_f32(xf, out)
    assert _rel(out, xf.double().sum(0)) < 1e-5
    # l2norm
    h = torch.randn(100, 256, device="cuda", generator=g)
    y, inv = ops.l2norm_fwd(h)
    hr = h.double().requires_grad_(True)
    yr = hr / hr.norm(dim=-1, keepdim=True)
    assert (y.double() - yr).abs().max() < 4e-3
    dy = torch.randn(100, 256, device="cuda", generator=g)
    (yr * dy.double()).sum().backward()
    dx = ops.l2norm_bwd(h, inv, dy)
    assert _rel(dx.float(), hr.grad) < 5e-3
    # weight norm
    v = torch.randn(4096, 256, device="cuda", generator=g) * 0.02
    gg = 1 + 0.05 * torch.randn(4096, 1, device="cuda", generator=g)
    w, winv = ops.weightnorm_fwd(v, gg)
    vr = v.double().requires_grad_(True)
    gr = gg.double().requires_grad_(True)
    wr = vr * (gr / vr.norm(dim=1, keepdim=True))
    assert _rel(w.float(), wr) < 4e-3
    dw = torch.randn(4096, 256, device="cuda", generator=g)
    (wr * dw.double()).sum().backward()
    dv, dg = ops.weightnorm_bwd(dw, v, gg, winv)
    assert _rel(dv, vr.grad) < 1e-4 and _rel(dg, gr.grad) < 1e-4


def test_multi_tensor_cast_ema(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    srcs = [torch.randn(n, device="cuda", generator=g) for n in (7, 70000, 384 * 1152, 1)]
    dsts = [torch.empty(t.shape, dtype=torch.bfloat16, device="cuda") for t in srcs]
    tab = ops.ChunkTable()
    t, n = tab.get(srcs, dsts, 2)
    ops.multi_tensor(ops.MT_CAST_BF16, t, n)
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s.to(torch.bfloat16))
    teach = [torch.randn_like(s) for s in srcs]
    want = [0.99 * a + 0.01 * b for a, b in zip(teach, srcs)]
    tab2 = ops.ChunkTable()
    t, n = tab2.get(srcs, teach, 4)
    ops.multi_tensor(ops.MT_EMA, t, n, 0.99, 0.01)
    for a, b in zip(teach, want):
        assert (a - b).abs().max() < 1e-6


# ---------------------------------------------------------------------------------------------------------
# loss kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,K", [(26, 65536), (7, 1024), (3, 4096)])
def test_dino_ce(ops, R, K):
    import ccd_oracle as O
    g = torch.Generator(device="cuda").manual_seed(R)
    zs = torch.randn(2 * R, K, device="cuda", generator=g) * 0.3
    zt = torch.randn(2 * R, K, device="cuda", generator=g) * 0.3
    c = torch.randn(1, K, device="cuda", generator=g) * 0.05
    loss, stats = ops.dino_ce_fwd(zs, zt, c, 0.1, 0.04)
    x = zs.double().requires_grad_(True)
    ref = O.dino_distill_loss(x, zt.double(), c.double(), 0.04, 0.1)
    assert abs(loss.item() - ref.item()) / abs(ref.item()) < 1e-5
    (ref * 1.7).backward()
    gs = torch.tensor([1.7], device="cuda")
    dz = ops.dino_ce_bwd(zs, zt, c, stats, gs, 0.1, 0.04)
    assert _rel(dz.float(), x.grad) < 5e-3       # bf16 output


def test_seg_ce(ops):
    import ccd_oracle as O
    g = torch.Generator(device="cuda").manual_seed(4)
    lg = torch.randn(6, 2, 32, 128, device="cuda", generator=g) * 2
    gt = (torch.rand(6, 32, 128, device="cuda", generator=g) > 0.6).float()
    loss = ops.seg_ce_fwd(lg, gt)
    x = lg.double().requires_grad_(True)
    ref = O.seg_loss(x, gt)
    assert abs(loss.item() - ref.item()) < 1e-5
    ref.backward()
    d = ops.seg_ce_bwd(lg, gt, None)
    assert _rel(d, x.grad) < 1e-4


def test_center_update(ops):
    import ccd_oracle as O
    g = torch.Generator(device="cuda").manual_seed(9)
    zt = torch.randn(52, 65536, device="cuda", generator=g)
    c = torch.randn(1, 65536, device="cuda", generator=g) * 0.01
    want = O.updated_center(c.clone(), zt)
    ops.center_update(c, zt, 1, 0.9)
    assert (c - want).abs().max() < 1e-5


# ---------------------------------------------------------------------------------------------------------
# character segments (bit exact)
# ---------------------------------------------------------------------------------------------------------
def test_ccl_against_reference_golden(ops):
    z = np.load(os.path.join(GOLD, "ccl_cases.npz"))
    masks = torch.tensor(z["masks"]).float().cuda()
    bits, compact, ncomp = ops.ccl_label(masks, 0, masks.shape[0], want_compact=True)
    assert np.array_equal(compact.cpu().numpy(), z["compact"])
    want_bits = np.where(z["compact"] > 0, 1 << (z["compact"].astype(np.int64) - 1), 0).astype(np.int32)
    assert np.array_equal(bits.cpu().numpy(), want_bits)
    assert np.array_equal(ncomp.cpu().numpy(), z["compact"].reshape(len(z["compact"]), -1).max(1))


def test_ccl_random_vs_oracle(ops):
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    masks = torch.cat([S.random_masks(24, seed=101), S.random_masks(8, seed=102, blobs=False, density=0.6)])
    _, compact, _ = ops.ccl_label(masks.cuda(), 0, masks.shape[0], want_compact=True)
    for i, m in enumerate(masks.numpy()):
        assert np.array_equal(compact[i].cpu().numpy(), O.label_cluster(m)[1]), i


def test_ccl_from_seg_logits(ops):
    import ccd_oracle as O
    g = torch.Generator().manual_seed(0)
    base = torch.nn.functional.avg_pool2d(torch.randn(4, 1, 32, 128, generator=g), 5, 1, 2)[:, 0]
    logits = torch.stack([-base, base], 1).contiguous()
    _, compact, _ = ops.ccl_label(logits.cuda(), 1, 4, want_compact=True)
    for i in range(4):
        m = (torch.softmax(logits[i], 0)[1] > 0.5).int().numpy()
        assert np.array_equal(compact[i].cpu().numpy(), O.label_cluster(m)[1])


def test_warp_and_dense(ops):
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    _, masks, metrics = S.make_batch(12, seed=7)
    masks, metrics = masks.cuda(), metrics.cuda()
    bits, _, _ = ops.ccl_label(masks, 0, 12)
    dense = ops.bits_to_dense(bits)
    want = torch.tensor(np.stack([O.label_cluster(m)[0] for m in masks.cpu().numpy()]).astype(np.float32)).cuda()
    assert torch.equal(dense, want)
    assert torch.equal(ops.dense_to_bits(dense), bits)
    wb = ops.warp_bits(bits, metrics)
    ref = O.warp_affine_binary(want.cpu(), metrics.cpu()).cuda()       # ATen CPU grid_sample as the checker
    got = ops.bits_to_dense(wb)
    mism = (got != ref).sum().item()
    assert mism == 0, f"{mism} warped cluster pixels differ"
    wm = ops.warp_mask(masks, metrics)
    refm = O.warp_affine_binary(masks.cpu().unsqueeze(1), metrics.cpu()).squeeze(1).cuda()
    assert (wm != refm).sum().item() == 0


@pytest.mark.parametrize("E", [192, 384])
def test_char_pool(ops, E):
    import ccd_oracle as O
    from ccd_b200 import synthetic as S
    B = 6
    _, masks, metrics = S.make_batch(B, seed=3)
    masks[1] = 0                                    # empty mask -> 4 zero rows (clamp to 3)
    masks[2, :, 64:] = 0
    masks, metrics = masks.cuda(), metrics.cuda()
    bits1, _, _ = ops.ccl_label(masks, 0, B)
    bits = torch.cat([bits1, ops.warp_bits(bits1, metrics)])
    tot4, cnt, offs, new_index = ops.char_plan(bits)
    R = int(offs[-1].item())
    g = torch.Generator(device="cuda").manual_seed(E)
    tokens = torch.randn(2 * B * 256, E, device="cuda", generator=g)
    rows = ops.char_pool_fwd(tokens, bits, tot4, cnt, offs, R)
    clusters = ops.bits_to_dense(bits)
    tr = tokens.double().reshape(2 * B, 256, E).requires_grad_(True)
    pooled, index = O.char_pool(tr, clusters.double())
    want_rows, want_index = O.ragged_select(pooled, index)
    assert torch.equal(new_index.bool(), want_index)
    assert rows.shape == want_rows.shape
    assert (rows.double() - want_rows).abs().max() < 1e-5
    drows = torch.randn(2 * R, E, device="cuda", generator=g)
    (want_rows * drows.double()).sum().backward()
    dtok = ops.char_pool_bwd(drows, bits, tot4, cnt, offs, E)
    assert (dtok.double().reshape(2 * B, 256, E) - tr.grad).abs().max() < 1e-5


def test_patch_im2col(ops):
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(3, 3, 32, 128, device="cuda", generator=g)
    cols = ops.patch_im2col(x)
    want = torch.nn.functional.unfold(x, 4, stride=4).transpose(1, 2).reshape(3 * 256, 48)
    assert torch.equal(cols[:, :48], want.to(torch.bfloat16))
    assert (cols[:, 48:] == 0).all()


def test_multi_tensor_clip(ops):
    import ccd_oracle as O
    from ccd_b200 import train_utils as U
    g = torch.Generator(device="cuda").manual_seed(12)
    lin = torch.nn.Sequential(torch.nn.Linear(300, 400), torch.nn.LayerNorm(400), torch.nn.Linear(400, 3)).cuda()
    scales = [10.0, 0.001, 0.2, 5.0, 0.5, 0.01]
    for p, s in zip(lin.parameters(), scales):
        p.grad = torch.randn(p.shape, device="cuda", generator=g) * s
    want = O.clip_per_parameter({n: p.grad.clone() for n, p in lin.named_parameters()}, 3.0)
    norms0 = [p.grad.norm().item() for p in lin.parameters()]
    norms = U.clip_gradients(lin, 3.0)
    for (n, p), n0, n1 in zip(lin.named_parameters(), norms0, norms.tolist()):
        assert abs(n0 - n1) < 1e-3 * n0
        assert (p.grad - want[n]).abs().max() < 1e-5 * max(1.0, want[n].abs().max().item())


def test_fused_adamw_clip_ema(ops):
    """optim.AdamW.step(clip_grad, ema) == clip_gradients (oracle) -> torch.optim.AdamW -> EMA, incl. parameters that skip a
    step (frozen last layer: different step counts), a weight-decay-free group, EMA-only pairs and bf16 copies."""
    import ccd_oracle as O
    from ccd_b200.optim import AdamW

    class Holder:            # stands in for TeacherEMA + the modules that own bf16 operand copies
        pass

    g = torch.Generator(device="cuda").manual_seed(21)
    shapes = [(300, 130), (70001,), (384, 1152), (2,), (17, 3)]
    mine = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    frozen = torch.nn.Parameter(torch.randn(33, device="cuda", generator=g), requires_grad=False)   # EMA only (weight_g)
    teach = [torch.randn(s, device="cuda", generator=g) for s in shapes] + [torch.randn(33, device="cuda", generator=g)]
    teach_ref = [t.clone() for t in teach]
    bf_s = torch.empty(shapes[2], dtype=torch.bfloat16, device="cuda")
    bf_t = torch.empty(shapes[2], dtype=torch.bfloat16, device="cuda")

    class Mod:
        def __init__(self, p, b):
            self.p, self.b, self.ver = p, b, None
        def bf16_copies(self):
            return [(self.p, self.b)]
        def bf16_is_fresh(self):
            return True
        def bf16_mark_fresh(self):
            self.ver = self.p._version

    ema = Holder()
    ema.pairs = list(zip(mine + [frozen], teach))
    ema.bf16_modules = [Mod(mine[2], bf_s), Mod(teach[2], bf_t)]
    groups = lambda ps: [{"params": [ps[0], ps[2], ps[4]]}, {"params": [ps[1], ps[3]], "weight_decay": 0.0}]
    opt = AdamW(groups(mine))
    opt_ref = torch.optim.AdamW(groups(ref))
    clip, mom = 3.0, 0.99
    for it in range(4):
        lr, wd = 1e-3 * (it + 1), 0.04 + 0.01 * it
        for o in (opt, opt_ref):
            for i, gr in enumerate(o.param_groups):
                gr["lr"] = lr
                if i == 0:
                    gr["weight_decay"] = wd
        grads = [torch.randn(s, device="cuda", generator=g) * sc for s, sc in zip(shapes, (0.05, 0.001, 0.5, 3.0, 0.2))]
        skip = {4} if it == 1 else set()            # parameter 4 has no gradient on step 1 -> its step count lags
        clipped = O.clip_per_parameter({str(i): gr.clone() for i, gr in enumerate(grads)}, clip)
        for i, (p, r) in enumerate(zip(mine, ref)):
            p.grad = None if i in skip else grads[i].clone()
            r.grad = None if i in skip else clipped[str(i)].clone()
        opt.step(clip_grad=clip, ema=ema, ema_momentum=mom)
        opt_ref.step()
        with torch.no_grad():
            for t, s in zip(teach_ref, ref + [frozen]):
                t.mul_(mom).add_((1 - mom) * s.detach())
        for p, r in zip(mine, ref):
            assert (p - r).abs().max().item() < 2e-6 * max(1.0, r.abs().max().item())
        for t, tr in zip(teach, teach_ref):
            assert (t - tr).abs().max().item() < 2e-6 * max(1.0, tr.abs().max().item())
        assert torch.equal(bf_s, mine[2].detach().to(torch.bfloat16)) and torch.equal(bf_t, teach[2].to(torch.bfloat16))
        assert ema.bf16_modules[0].ver == mine[2]._version
    sd = opt.state_dict()
    assert torch.is_tensor(sd["state"][0]["step"]) and float(sd["state"][0]["step"]) == 4.0 and float(sd["state"][2]["step"]) == 3.0
    opt_ref.load_state_dict(sd)                     # interchange with torch.optim.AdamW
    opt.load_state_dict(opt_ref.state_dict())
    assert opt.state[mine[4]]["step"] == 3


def test_gelu_epilogue_accuracy(ops):
    """Fused GELU / GELU' epilogues (packed polynomial + MUFU.TANH) against the exact erf form over x in [-12, 12]:
    absolute error beyond the bf16 rounding of the stored result stays below 1.5e-3 (GELU) / 3e-3 (GELU')."""
    import json
    M, N, K = 4096, 256, 64
    x = torch.linspace(-12.0, 12.0, M, device="cuda").to(torch.bfloat16)
    A = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    A[:, 0] = x
    B = torch.zeros(N, K, dtype=torch.bfloat16, device="cuda")
    B[:, 0] = 1.0
    act = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_GELU, None, None, act)
    xd = x.double()
    exact = 0.5 * xd * (1.0 + torch.erf(xd / math.sqrt(2.0)))
    err = (act[:, 0].double() - exact).abs() - exact.abs() * 2.0 ** -8
    assert torch.equal(act[:, 0], act[:, N - 1])
    A[:, 0] = 1.0
    aux = x.view(M, 1).expand(M, N).contiguous()
    dout = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M, N, K, 0, 0, ops.EPI_DGELU, None, dout, None, aux)
    dexact = 0.5 * (1.0 + torch.erf(xd / math.sqrt(2.0))) + xd * torch.exp(-0.5 * xd * xd) / math.sqrt(2.0 * math.pi)
    derr = (dout[:, 0].double() - dexact).abs() - dexact.abs() * 2.0 ** -8
    rec = {"gelu_max_excess_abs_err": err.max().item(), "gelu_argmax_x": x[err.argmax()].item(),
           "dgelu_max_excess_abs_err": derr.max().item(), "dgelu_argmax_x": x[derr.argmax()].item()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gelu_accuracy.json", "w") as f:
        json.dump(rec, f)
    assert err.max().item() < 1.5e-3 and derr.max().item() < 3e-3, rec


# ---------------------------------------------------------------------------------------------------------
# text-mask generation (first slice)
# ---------------------------------------------------------------------------------------------------------
def test_kmeans_mask_bit_exact_vs_oracle_and_reference_golden(ops):
    """ccd_kmeans_mask against the numpy oracle (bit-exact: integer histogram + the same float64 objective) on the golden crops,
    odd sizes, constant and two-level images; against the reference's own masks (tests/golden/kmeans_masks.npz) with the
    oracle's criteria; and end to end into the component labelling."""
    import os
    import mask_oracle as MO
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_masks.npz"))
    imgs = z["images"]
    got = ops.kmeans_mask(torch.from_numpy(imgs).cuda()).cpu().numpy().astype(np.uint8)
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i], MO.cluster_pixels(im)), i
    exact = sum(np.array_equal(g, w) for g, w in zip(got, z["masks"]))
    assert exact >= 0.85 * len(imgs)
    assert min((g == w).mean() for g, w in zip(got, z["masks"])) >= 0.985
    # other sizes + degenerate images
    for h, w, seed in ((48, 160, 3), (17, 33, 4), (64, 256, 5)):
        ims = MO.synthetic_text_crops(5, h=h, w=w, seed=seed)
        ims[0] = 93                                            # constant -> all zero
        ims[1] = 200
        ims[1, 2:h - 2, 5:9] = 15                              # two levels, dark text
        out = ops.kmeans_mask(torch.from_numpy(ims).cuda()).cpu().numpy().astype(np.uint8)
        for i in range(len(ims)):
            assert np.array_equal(out[i], MO.cluster_pixels(ims[i])), (h, w, i)
        assert out[0].sum() == 0 and out[1].sum() == (h - 4) * 4
    # the masks feed the component labelling directly (same result as labelling the oracle's masks)
    import ccd_oracle as O
    m = ops.kmeans_mask(torch.from_numpy(imgs[:8]).cuda())
    bits, compact, ncomp = ops.ccl_label(m, 0, 8, want_compact=True)
    for i in range(8):
        _, want = O.label_cluster(MO.cluster_pixels(imgs[i]).astype(np.float32))
        assert np.array_equal(compact[i].cpu().numpy(), want), i


def test_affine_theta_matches_the_dataset_algebra(ops):
    """ccd_affine_theta against the line-for-line numpy restatement of datasetsupervised_kmeans.py:63-71, and the end-to-end meaning
    of the result: warping with theta through F.affine_grid / grid_sample reproduces the pixel-space warp the matrix describes."""
    import mask_oracle as MO
    g = np.random.default_rng(5)
    n = 64
    m_inv = np.tile(np.eye(3), (n, 1, 1))
    hw = np.stack([g.integers(24, 200, n), g.integers(60, 600, n)], 1).astype(np.int32)
    for i in range(1, n):                                    # sample 0 stays the identity
        a, sh = g.uniform(-0.17, 0.17), g.uniform(-0.6, 0.6)
        sx, sy = g.uniform(0.6, 1.1), g.uniform(0.6, 1.1)
        fwd = np.array([[sx * np.cos(a), -sy * np.sin(a + sh), g.uniform(-5, 5)], [sx * np.sin(a), sy * np.cos(a + sh), g.uniform(-2, 2)], [0, 0, 1]])
        m_inv[i] = np.linalg.inv(fwd)
    got = ops.affine_theta(torch.from_numpy(m_inv).cuda(), torch.from_numpy(hw).cuda()).cpu().numpy()
    for i in range(n):
        want = MO.affine_theta(m_inv[i], int(hw[i, 0]), int(hw[i, 1]))
        assert np.allclose(got[i], want, rtol=2e-6, atol=2e-6), (i, got[i], want)
    assert np.allclose(got[0], np.eye(3), atol=1e-7)
    # meaning: theta maps normalised output coordinates to normalised input coordinates of the 32 x 128 crop
    i = 7
    th = torch.from_numpy(got[i:i + 1, :2, :])
    grid = torch.nn.functional.affine_grid(th, (1, 1, 32, 128), align_corners=True)       # N = [[2/(w-1),0,-1],...] is the align_corners=True map
    ys, xs = 11, 57
    ws, hs = hw[i, 1] / 128.0, hw[i, 0] / 32.0
    src_px = m_inv[i] @ np.array([xs * ws, ys * hs, 1.0])
    want_norm = np.array([src_px[0] / ws * 2 / 127 - 1, src_px[1] / hs * 2 / 31 - 1])
    assert np.allclose(grid[0, ys, xs].numpy(), want_norm, atol=1e-4)
