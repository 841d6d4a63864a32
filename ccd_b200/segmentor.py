"""Segmentation head (drop-in for Dino/modules/segmentor.py:37-95, same parameter / buffer names and shapes) on the
sm_100a kernels.

  3 x [conv3x3 E->128 (no bias), BN, ReLU, conv1x1 128->64 (no bias), BN, ReLU] on the three norm_seg taps [N,E,8,32]
  -> cat 192 -> ConvT 4x4 s2 (192->128) BN ReLU @16x64 -> ConvT 4x4 s2 (128->128) BN ReLU @32x128 -> conv3x3 128->2

Every convolution is an implicit GEMM on the persistent tcgen05 kernel (`ccd_conv_gemm`): activations are NHWC bf16,
the A (or, for weight gradients, B) operand is gathered by 5-D TMA boxes shifted per filter tap, zero fill = padding;
the stride-2 transposed convolutions are four 2x2-tap GEMMs (one per output parity) whose epilogue scatters rows to the
interleaved output, and their gradients gather the parity planes of the upsampled gradient through a {2*C, W, 2, H, N}
view.  BatchNorm (training statistics, SyncBatchNorm when converted, train.py:96-98) + ReLU are row-streaming kernels.
`conv_mla` is constructed but never called, exactly like the reference (its parameters never receive gradients).
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


def _cbr(cin, cout, k, pad):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class Conv_MLA(nn.Module):          # segmentor.py:6-35 -- parameters only (unused by forward)
    def __init__(self, in_channels=1024, mla_channels=256):
        super().__init__()
        self.mla_p2_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p3_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p4_1x1 = _cbr(in_channels, mla_channels, 1, 0)
        self.mla_p2 = _cbr(mla_channels, mla_channels, 3, 1)
        self.mla_p3 = _cbr(mla_channels, mla_channels, 3, 1)
        self.mla_p4 = _cbr(mla_channels, mla_channels, 3, 1)


class MLAHead(nn.Module):           # segmentor.py:37-70 (parameter container; SegHead.forward drives the kernels)
    def __init__(self, in_channels=384, mla_channels=128, mlahead_channels=64):
        super().__init__()

        def branch():
            return nn.Sequential(nn.Conv2d(in_channels, mla_channels, 3, padding=1, bias=False), nn.BatchNorm2d(mla_channels),
                                 nn.ReLU(), nn.Conv2d(mla_channels, mlahead_channels, 1, bias=False),
                                 nn.BatchNorm2d(mlahead_channels), nn.ReLU())
        self.head2, self.head3, self.head4 = branch(), branch(), branch()


# filter taps ------------------------------------------------------------------------------------------------
CONV3_FWD_TAPS = [(ky - 1, kx - 1, 0, 0) for ky in range(3) for kx in range(3)]          # input pixel (y+ky-1, x+kx-1)
CONV3_DGRAD_TAPS = [(1 - ky, 1 - kx, 0, 0) for ky in range(3) for kx in range(3)]        # dY pixel (y-(ky-1), x-(kx-1))
# ConvTranspose2d(k=4, s=2, p=1): out[2a+py] gets in[a + d] * W[ky] for (d, ky) in
_CT_FWD = {0: ((0, 1), (-1, 3)), 1: ((0, 2), (1, 0))}
# its data gradient: in[a] gets dOut[2(a+d)+p] * W[ky] for ky -> (d, p)
_CT_BWD = {0: (-1, 1), 1: (0, 0), 2: (0, 1), 3: (1, 0)}


def convt_fwd_taps(py, px):
    """-> (taps, [(ky, kx)]) of the 2x2 sub-filter producing output parity (py, px)."""
    taps, kk = [], []
    for dy, ky in _CT_FWD[py]:
        for dx, kx in _CT_FWD[px]:
            taps.append((dy, dx, 0, 0))
            kk.append((ky, kx))
    return taps, kk


def convt_bwd_taps(cout):
    """16 taps over the {2*cout, W, 2, H, N} parity view of the upsampled gradient; tap index = ky*4 + kx."""
    taps = []
    for ky in range(4):
        for kx in range(4):
            dy, py = _CT_BWD[ky]
            dx, px = _CT_BWD[kx]
            taps.append((dy, dx, py, px * cout))
    return taps


class SegHead(nn.Module):           # segmentor.py:73-95
    def __init__(self, in_channels=384, mla_channels=128, mlahead_channels=64, num_classes=2, **kwargs):
        super().__init__()
        if mla_channels != 128 or mlahead_channels != 64 or num_classes != 2 or in_channels % 64:
            raise NotImplementedError("ccd_b200.SegHead implements the CCD configuration (128 / 64 channels, 2 classes)")
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.conv_mla = Conv_MLA(in_channels, mla_channels)
        self.mlahead = MLAHead(in_channels=in_channels, mla_channels=mla_channels, mlahead_channels=mlahead_channels)
        self.unpool1 = nn.Sequential(nn.ConvTranspose2d(192, 128, (4, 4), (2, 2), (1, 1)), nn.BatchNorm2d(128), nn.ReLU(True))
        self.unpool2 = nn.Sequential(nn.ConvTranspose2d(128, 128, (4, 4), (2, 2), (1, 1)), nn.BatchNorm2d(128), nn.ReLU(True))
        self.cls = nn.Conv2d(128, self.num_classes, 3, padding=1)

    def _bn_layers(self):
        h = self.mlahead
        return [h.head2[1], h.head2[4], h.head3[1], h.head3[4], h.head4[1], h.head4[4], self.unpool1[1], self.unpool2[1]]

    def forward(self, inputs):
        if not inputs[0].is_cuda:
            raise RuntimeError("ccd_b200.SegHead runs on CUDA (sm_100a) only; there is no CPU path")
        h = self.mlahead
        params = []
        for br in (h.head2, h.head3, h.head4):
            params += [br[0].weight, br[1].weight, br[1].bias, br[3].weight, br[4].weight, br[4].bias]
        params += [self.unpool1[0].weight, self.unpool1[0].bias, self.unpool1[1].weight, self.unpool1[1].bias,
                   self.unpool2[0].weight, self.unpool2[0].bias, self.unpool2[1].weight, self.unpool2[1].bias,
                   self.cls.weight, self.cls.bias]
        # taps arrive as [N,E,8,32] views of NHWC storage (encoder.VisionTransformer.forward): recover [N*256, E]
        taps = [t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]) for t in inputs]
        return SegHeadFn.apply(self, *taps, *params)


def _sync(bn):
    return isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _bn_forward_group(items, training):
    """BatchNorm2d / SyncBatchNorm (training: batch statistics, running-stat update) + ReLU for several INDEPENDENT layers
    items = [(bn, z, ldz, M, C, y, ldy)]; under SyncBatchNorm their statistics travel in ONE all-reduce.
    Returns [(mean, rstd, count)]."""
    out = []
    if training:
        sums = [ops.bn_stats(z, ldz, M, C) for (_, z, ldz, M, C, _, _) in items]
        world = 1
        if _sync(items[0][0]):
            world = dist.get_world_size()
            if len(sums) > 1:
                flat = torch.cat(sums)
                dist.all_reduce(flat)
                sums = list(flat.split([t.numel() for t in sums]))
            else:
                dist.all_reduce(sums[0])
        for (bn, z, ldz, M, C, y, ldy), sm in zip(items, sums):
            # every rank holds the same number of rows: the reference's loader is built with drop_last=True (train.py:436-444)
            count = float(M) * world
            mom = 0.1 if bn.momentum is None else bn.momentum
            mean, rstd = ops.bn_finalize(sm, count, bn.eps, mom, bn.running_mean if bn.track_running_stats else None,
                                         bn.running_var if bn.track_running_stats else None, C)
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            out.append((mean, rstd, count))
    else:
        # eval: running statistics are constants of the layer -> count = inf marks "no batch-statistic terms" for the backward
        for (bn, z, ldz, M, C, y, ldy) in items:
            out.append((bn.running_mean.float(), torch.rsqrt(bn.running_var.float() + bn.eps), float("inf")))
    for (bn, z, ldz, M, C, y, ldy), (mean, rstd, _) in zip(items, out):
        ops.bn_apply_relu(z, ldz, mean, rstd, bn.weight.detach(), bn.bias.detach(), y, ldy, M, C)
    return out


def _bn_forward(bn, z, ldz, M, C, y, ldy, training):
    return _bn_forward_group([(bn, z, ldz, M, C, y, ldy)], training)[0]


def _bn_backward_group(items):
    """items = [(bn, dy, lddy, z, ldz, mean, rstd, count, M, C)] of independent layers -> [(dz bf16 [M,C], dgamma, dbeta)]."""
    sums = [ops.bn_bwd_reduce(dy, lddy, z, ldz, mean, rstd, bn.weight.detach(), bn.bias.detach(), M, C)
            for (bn, dy, lddy, z, ldz, mean, rstd, count, M, C) in items]
    # local sums = parameter gradients (dgamma, dbeta); DDP averages them across ranks
    pgrads = [(sm[it[9]:].clone(), sm[:it[9]].clone()) for sm, it in zip(sums, items)]
    # eval-mode forward (running statistics, count = inf): dz = dy * gamma * rstd under the ReLU mask -- no reduction terms
    # (1/count = 0 zeroes them in bn_bwd_apply) and no SyncBN exchange
    frozen = items[0][7] == float("inf")
    if _sync(items[0][0]) and not frozen:
        if len(sums) > 1:
            flat = torch.cat(sums)
            dist.all_reduce(flat)
            sums = list(flat.split([t.numel() for t in sums]))
        else:
            dist.all_reduce(sums[0])
    out = []
    for (bn, dy, lddy, z, ldz, mean, rstd, count, M, C), sm, (dgamma, dbeta) in zip(items, sums, pgrads):
        dz = ops.bn_bwd_apply(dy, lddy, z, ldz, mean, rstd, bn.weight.detach(), bn.bias.detach(), sm, 1.0 / count, M, C)
        out.append((dz, dgamma, dbeta))
    return out


def _bn_backward(bn, dy, lddy, z, ldz, mean, rstd, count, M, C):
    return _bn_backward_group([(bn, dy, lddy, z, ldz, mean, rstd, count, M, C)])[0]


class SegHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, t0, t1, t2, *params):
        dev = t0.device
        E = mod.in_channels
        T = t0.shape[0]
        n = T // 256
        M1, M2 = n * 1024, n * 4096
        b16 = dict(dtype=torch.bfloat16, device=dev)
        training = mod.training
        h = mod.mlahead
        branches = (h.head2, h.head3, h.head4)
        saved = {"n": n, "bn": []}
        cat = torch.empty(T, 192, **b16)
        xs, zas, acts, zbs = [], [], [], []
        for i, (br, tap) in enumerate(zip(branches, (t0, t1, t2))):
            w0 = params[6 * i].detach()
            x = ops.cast_bf16(tap.contiguous().float()) if tap.dtype != torch.bfloat16 else tap.contiguous()
            wf = w0.permute(0, 2, 3, 1).reshape(128, 9 * E).to(torch.bfloat16).contiguous()         # [o][tap][c]
            za = torch.empty(T, 128, **b16)
            ops.conv_gemm(x, wf, T, 128, 9 * E, ops.EPI_BF16, None, za, 128, 1, 1, 8, 32, E, 1, n, CONV3_FWD_TAPS, E)
            xs.append(x); zas.append(za); acts.append(torch.empty(T, 128, **b16))
        # the three branches are independent: their BatchNorm statistics share one (Sync)BN exchange per level
        st_as = _bn_forward_group([(br[1], zas[i], 128, T, 128, acts[i], 128) for i, br in enumerate(branches)], training)
        for i, br in enumerate(branches):
            w3 = params[6 * i + 3].detach()
            zb = torch.empty(T, 64, **b16)
            ops.linear_fwd(acts[i], w3.reshape(64, 128).to(torch.bfloat16).contiguous(), None, ops.EPI_BF16, zb)
            zbs.append(zb)
        st_bs = _bn_forward_group([(br[4], zbs[i], 64, T, 64, cat[:, 64 * i:], 192) for i, br in enumerate(branches)], training)
        for i in range(3):
            saved["bn"] += [st_as[i], st_bs[i]]
        # unpool1: ConvTranspose2d(192 -> 128) 8x32 -> 16x64, four parity GEMMs
        wu1, bu1 = params[18].detach(), params[19].detach()
        zu1 = torch.empty(M1, 128, **b16)
        for py in range(2):
            for px in range(2):
                taps, kk = convt_fwd_taps(py, px)
                wpar = torch.stack([wu1[:, :, ky, kx] for ky, kx in kk], 0).permute(2, 0, 1).reshape(128, 4 * 192)
                ops.conv_gemm(cat, wpar.to(torch.bfloat16).contiguous(), T, 128, 4 * 192, ops.EPI_BF16, bu1.float(), zu1, 128, 1, 1,
                              8, 32, 192, 1, n, taps, 192, rowmap=1, py=py, px=px)
        u1 = torch.empty(M1, 128, **b16)
        st_u1 = _bn_forward(mod.unpool1[1], zu1, 128, M1, 128, u1, 128, training)
        # unpool2: ConvTranspose2d(128 -> 128) 16x64 -> 32x128
        wu2, bu2 = params[22].detach(), params[23].detach()
        zu2 = torch.empty(M2, 128, **b16)
        for py in range(2):
            for px in range(2):
                taps, kk = convt_fwd_taps(py, px)
                wpar = torch.stack([wu2[:, :, ky, kx] for ky, kx in kk], 0).permute(2, 0, 1).reshape(128, 4 * 128)
                ops.conv_gemm(u1, wpar.to(torch.bfloat16).contiguous(), M1, 128, 4 * 128, ops.EPI_BF16, bu2.float(), zu2, 128, 1, 1,
                              16, 64, 128, 1, n, taps, 128, rowmap=1, py=py, px=px)
        u2 = torch.empty(M2, 128, **b16)
        st_u2 = _bn_forward(mod.unpool2[1], zu2, 128, M2, 128, u2, 128, training)
        # cls: conv3x3 128 -> 2, CUDA-core kernel (fp32 weights, NCHW fp32 logits)
        logits = ops.seg_cls_fwd(u2, params[26].detach().float().contiguous(), params[27].detach().float().contiguous(), n)
        if any(ctx.needs_input_grad):
            saved["bn"] += [st_u1, st_u2]
            ctx.mod, ctx.E, ctx.params = mod, E, params
            ctx.saved = (saved, xs, zas, acts, zbs, cat, zu1, u1, zu2, u2)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        mod, E, params = ctx.mod, ctx.E, ctx.params
        saved, xs, zas, acts, zbs, cat, zu1, u1, zu2, u2 = ctx.saved
        n = saved["n"]
        T, M1, M2 = n * 256, n * 1024, n * 4096
        dev = cat.device
        b16 = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        h = mod.mlahead
        branches = (h.head2, h.head3, h.head4)
        grads = [None] * len(params)
        st = saved["bn"]
        # ---- cls ----
        wc = params[26].detach().float().contiguous()
        dl = d_logits.float().contiguous()
        grads[26], grads[27] = ops.seg_cls_wgrad(u2, dl, n)
        du2 = ops.seg_cls_dgrad(dl, wc, n)
        # ---- unpool2 ----
        dz2, grads[24], grads[25] = _bn_backward(mod.unpool2[1], du2, 128, zu2, 128, *st[7], M2, 128)
        grads[23] = ops.colsum_bf16(dz2, torch.zeros(128, **f32))
        wu2 = params[22].detach()
        tb = convt_bwd_taps(128)
        gw2 = torch.zeros(128, 16 * 128, **f32)                                  # [cin][tap][o]
        ops.conv_gemm(dz2, u1, 128, 16 * 128, M1, ops.EPI_F32, None, gw2, 16 * 128, ops.wgrad_splits(128, 16 * 128, M1), 2, 16, 64, 256, 2,
                      n, tb, 128)
        grads[22] = gw2.reshape(128, 4, 4, 128).permute(0, 3, 1, 2).contiguous()
        wd2 = wu2.permute(0, 2, 3, 1).reshape(128, 16 * 128).to(torch.bfloat16).contiguous()      # [cin][tap][o]
        du1 = torch.empty(M1, 128, **b16)
        ops.conv_gemm(dz2, wd2, M1, 128, 16 * 128, ops.EPI_BF16, None, du1, 128, 1, 1, 16, 64, 256, 2, n, tb, 128)
        # ---- unpool1 ----
        dz1, grads[20], grads[21] = _bn_backward(mod.unpool1[1], du1, 128, zu1, 128, *st[6], M1, 128)
        grads[19] = ops.colsum_bf16(dz1, torch.zeros(128, **f32))
        wu1 = params[18].detach()
        gw1 = torch.zeros(192, 16 * 128, **f32)
        ops.conv_gemm(dz1, cat, 192, 16 * 128, T, ops.EPI_F32, None, gw1, 16 * 128, ops.wgrad_splits(192, 16 * 128, T), 2, 8, 32, 256, 2, n,
                      tb, 128)
        grads[18] = gw1.reshape(192, 4, 4, 128).permute(0, 3, 1, 2).contiguous()
        wd1 = wu1.permute(0, 2, 3, 1).reshape(192, 16 * 128).to(torch.bfloat16).contiguous()
        dcat = torch.empty(T, 192, **b16)
        ops.conv_gemm(dz1, wd1, T, 192, 16 * 128, ops.EPI_BF16, None, dcat, 192, 1, 1, 8, 32, 256, 2, n, tb, 128)
        # ---- the three branches ----
        d_taps = []
        rb = _bn_backward_group([(br[4], dcat[:, 64 * i:], 192, zbs[i], 64, *st[2 * i + 1], T, 64) for i, br in enumerate(branches)])
        das = []
        for i, br in enumerate(branches):
            w3 = params[6 * i + 3].detach()
            dzb, grads[6 * i + 4], grads[6 * i + 5] = rb[i]
            gw3 = torch.zeros(64, 128, **f32)
            ops.linear_wgrad(dzb, acts[i], gw3)
            grads[6 * i + 3] = gw3.reshape(64, 128, 1, 1)
            da = torch.empty(T, 128, **b16)
            ops.linear_dgrad(dzb, w3.reshape(64, 128).to(torch.bfloat16).contiguous(), ops.EPI_BF16, da)
            das.append(da)
        ra = _bn_backward_group([(br[1], das[i], 128, zas[i], 128, *st[2 * i], T, 128) for i, br in enumerate(branches)])
        for i, br in enumerate(branches):
            w0 = params[6 * i].detach()
            dza, grads[6 * i + 1], grads[6 * i + 2] = ra[i]
            gw0 = torch.zeros(128, 9 * E, **f32)
            ops.conv_gemm(xs[i], dza, 128, 9 * E, T, ops.EPI_F32, None, gw0, 9 * E, ops.wgrad_splits(128, 9 * E, T), 2, 8, 32, E, 1, n,
                          CONV3_FWD_TAPS, E)
            grads[6 * i] = gw0.reshape(128, 3, 3, E).permute(0, 3, 1, 2).contiguous()
            wd0 = w0.permute(1, 2, 3, 0).reshape(E, 9 * 128).to(torch.bfloat16).contiguous()      # [c][tap][o]
            dt = torch.empty(T, E, **f32)
            ops.conv_gemm(dza, wd0, T, E, 9 * 128, ops.EPI_F32, None, dt, E, 1, 1, 8, 32, 128, 1, n, CONV3_DGRAD_TAPS, 128)
            d_taps.append(dt)
        return (None, *d_taps, *grads)
