"""DINO projection head (drop-in for DINOHead, Dino/modules/vision_transformer.py:294-328) on the sm_100a kernels.

  rows [2R,E] -> Linear E->2048 + GELU -> Linear 2048->2048 + GELU -> Linear 2048->256 -> L2 normalise
       -> weight-normed Linear 256 -> K (no bias)  = logits [2R,K] fp32

Forward/backward are one autograd.Function over the C ABI: tcgen05 GEMMs with bias/GELU/GELU' epilogues, the
weight-norm scale g/||v|| folded into the bf16 operand copy of the last layer, and (fast path) the bf16 dlogits
written by the distillation-loss backward consumed directly as the GEMM A operand.
"""
import weakref

import torch
import torch.nn as nn

from . import ops
from .encoder import trunc_normal_

# Hand-off of the bf16 dlogits from the distillation-loss backward to the head backward (no fp32 [2R,K] gradient is ever
# written).  HeadFn.forward registers the logits it produced (data_ptr -> weak reference to its autograd node; a dead node =
# a stale entry); loss.DinoCEFn.backward uses the hand-off ONLY for registered logits and returns an ordinary dense
# gradient for anything else (a foreign head, a clone / slice / scaled view, a leaf tensor in a unit test).
HEAD_LOGITS = {}          # data_ptr -> (weakref(ctx of HeadFn), shape)
DLOGITS_STASH = {}        # data_ptr -> (bf16 dlogits, data_ptr of the zero-stride placeholder autograd carries instead)


def head_logits_registered(t):
    ent = HEAD_LOGITS.get(t.data_ptr())
    return ent is not None and ent[0]() is not None and ent[1] == tuple(t.shape) and t.is_contiguous() \
        and t.dtype == torch.float32


class _LastLayer(nn.Module):
    """Parameter container with the names nn.utils.weight_norm produces (weight_g [K,1], weight_v [K,256])."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        v = torch.empty(out_dim, in_dim)
        nn.init.kaiming_uniform_(v, a=5 ** 0.5)          # nn.Linear default init, as in the reference (:313)
        self.weight_g = nn.Parameter(torch.ones(out_dim, 1))
        self.weight_v = nn.Parameter(v)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048, bottleneck_dim=256):
        super().__init__()
        if use_bn or nlayers != 3:
            raise NotImplementedError("ccd_b200.DINOHead implements the CCD configuration (nlayers=3, no BN)")
        self.mlp = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, hidden_dim), nn.GELU(),
                                 nn.Linear(hidden_dim, bottleneck_dim))
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                nn.init.constant_(m.bias, 0)
        self.last_layer = _LastLayer(bottleneck_dim, out_dim)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False
        self._cast = ops.ChunkTable()
        self._bf16 = None
        self._bf16_ver = None
        self._wn = None

    def _bf16_srcs(self):
        return [self.mlp[0].weight, self.mlp[2].weight, self.mlp[4].weight]

    def bf16_copies(self):
        return list(zip(self._bf16_srcs(), self._bf16)) if self._bf16 is not None else []

    # freshness rule: see VisionTransformer (encoder.py) -- every forward re-casts unless the fused optimizer vouched
    def bf16_mark_fresh(self):
        self._bf16_ver = tuple(p._version for p in self._bf16_srcs())

    def bf16_invalidate(self):
        self._bf16_ver = None

    def bf16_is_fresh(self):
        return self._bf16_ver is not None and self._bf16_ver == tuple(p._version for p in self._bf16_srcs())

    def _bf16_weights(self):
        params = self._bf16_srcs()
        srcs = [p.detach() for p in params]
        if self._bf16 is None or self._bf16[0].device != srcs[0].device:
            self._bf16 = [torch.empty(s.shape, dtype=torch.bfloat16, device=s.device) for s in srcs]
            self._wn = torch.empty(self.last_layer.weight_v.shape, dtype=torch.bfloat16, device=srcs[0].device)
            self._bf16_ver = None
        if not self.bf16_is_fresh():
            table, n = self._cast.get(srcs, self._bf16, 2)
            ops.multi_tensor(ops.MT_CAST_BF16, table, n)
        self._bf16_ver = None                                  # one-shot: consumed by this forward
        return self._bf16

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ccd_b200.DINOHead runs on CUDA (sm_100a) only; there is no CPU path")
        wb = self._bf16_weights()
        ll = self.last_layer
        return HeadFn.apply(x.contiguous().float(), self, wb, self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight,
                            self.mlp[2].bias, self.mlp[4].weight, self.mlp[4].bias, ll.weight_g, ll.weight_v)


class HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, wb, w1, b1, w2, b2, w3, b3, wg, wv):
        R2, E = x.shape
        dev = x.device
        K = wv.shape[0]
        b16 = dict(dtype=torch.bfloat16, device=dev)
        xb = ops.cast_bf16(x)
        hid = w1.shape[0]
        bott = w3.shape[0]
        p1, a1 = torch.empty(R2, hid, **b16), torch.empty(R2, hid, **b16)
        ops.linear_fwd(xb, wb[0], b1.detach(), ops.EPI_GELU, p1, a1)
        p2, a2 = torch.empty(R2, hid, **b16), torch.empty(R2, hid, **b16)
        ops.linear_fwd(a1, wb[1], b2.detach(), ops.EPI_GELU, p2, a2)
        h3 = torch.empty(R2, bott, dtype=torch.float32, device=dev)
        ops.linear_fwd(a2, wb[2], b3.detach(), ops.EPI_F32, h3)
        yn, inv = ops.l2norm_fwd(h3)
        wn, winv = ops.weightnorm_fwd(wv.detach(), wg.detach(), mod._wn)      # bf16 (g/||v||) v, never an fp32 W
        logits = torch.empty(R2, K, dtype=torch.float32, device=dev)
        ops.linear_fwd(yn, wn, None, ops.EPI_F32, logits)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(xb, p1, a1, p2, a2, h3, yn, inv, wn, winv, wg, wv)
            ctx.wb = wb
            ctx.logits_ptr = logits.data_ptr()
            for k in [k for k, (w, _) in HEAD_LOGITS.items() if w() is None]:        # forwards whose graph is gone
                HEAD_LOGITS.pop(k, None)
                DLOGITS_STASH.pop(k, None)
            HEAD_LOGITS[ctx.logits_ptr] = (weakref.ref(ctx), (R2, K))
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        xb, p1, a1, p2, a2, h3, yn, inv, wn, winv, wg, wv = ctx.saved_tensors
        wb = ctx.wb
        R2, E = xb.shape
        dev = xb.device
        K, bott = wv.shape
        hid = p1.shape[1]
        b16 = dict(dtype=torch.bfloat16, device=dev)
        HEAD_LOGITS.pop(ctx.logits_ptr, None)
        stash = DLOGITS_STASH.pop(ctx.logits_ptr, None)
        if stash is None:
            dz = ops.cast_bf16(d_logits.contiguous().float())      # generic path: any upstream gradient
        elif d_logits.data_ptr() == stash[1]:
            dz = stash[0]                                          # fast path: autograd carried only the placeholder
        else:
            # the logits also fed something else: autograd summed that gradient with the (all-zero) placeholder
            dz = ops.cast_bf16(d_logits.contiguous().float() + stash[0].float())
        # last layer
        dw = torch.zeros(K, bott, dtype=torch.float32, device=dev)
        ops.linear_wgrad(dz, yn, dw)
        dv, dg = ops.weightnorm_bwd(dw, wv.detach(), wg.detach(), winv)
        dyn = torch.empty(R2, bott, dtype=torch.float32, device=dev)
        ops.linear_dgrad(dz, wn, ops.EPI_F32, dyn)
        dh3 = ops.l2norm_bwd(h3, inv, dyn)
        # mlp[4]
        gw3 = torch.zeros(bott, hid, dtype=torch.float32, device=dev)
        gb3 = torch.zeros(bott, dtype=torch.float32, device=dev)
        ops.linear_wgrad(dh3, a2, gw3)
        ops.colsum_bf16(dh3, gb3)
        # mlp[2] / mlp[0]: their bias gradients (column sums of dp2 / dp1) come out of the GELU' epilogues
        gw2 = torch.zeros(hid, hid, dtype=torch.float32, device=dev)
        gb2 = torch.zeros(hid, dtype=torch.float32, device=dev)
        gw1 = torch.zeros(hid, E, dtype=torch.float32, device=dev)
        gb1 = torch.zeros(hid, dtype=torch.float32, device=dev)
        dp2 = torch.empty(R2, hid, **b16)
        ops.linear_dgrad(dh3, wb[2], ops.EPI_DGELU, dp2, p2, colsum=gb2)
        ops.linear_wgrad(dp2, a1, gw2)
        dp1 = torch.empty(R2, hid, **b16)
        ops.linear_dgrad(dp2, wb[1], ops.EPI_DGELU, dp1, p1, colsum=gb1)
        ops.linear_wgrad(dp1, xb, gw1)
        dx = torch.empty(R2, E, dtype=torch.float32, device=dev)
        ops.linear_dgrad(dp1, wb[0], ops.EPI_F32, dx)
        return (dx, None, None, gw1, gb1, gw2, gb2, gw3, gb3, dg if ctx.needs_input_grad[9] else None, dv)
