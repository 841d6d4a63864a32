"""ViT encoder of CCD (drop-in for Dino/modules/vision_transformer.py:134-291) on the sm_100a kernels.

Same constructor, attributes and parameter names/shapes as the reference `VisionTransformer` (so reference
checkpoints load and DDP / AdamW / EMA / get_params_groups keep working), but forward + backward run as one
autograd.Function that drives the C-ABI kernels:

  patch embed  : im2col -> tcgen05 GEMM with (bias + resampled pos-embed) epilogue          -> fp32 residual stream
  block        : LN (warp-shuffle) -> QKV GEMM -> fused MHSA (tcgen05/TMEM) -> proj GEMM + residual epilogue
                 -> LN -> fc1 GEMM + GELU epilogue -> fc2 GEMM + residual epilogue
  backward     : dgrad GEMMs read the weights MN-major, wgrad GEMMs read the activations MN-major (no transposes),
                 GELU' fused in the dgrad epilogue, LN backward fused with the residual-gradient add.
Parameters stay fp32 nn.Parameters; bf16 operand copies are refreshed by one multi-tensor cast per forward.
"""

import torch
import torch.nn as nn

from . import ops

GRID_TOKENS = 256
TAP_AFTER = (2, 4, 6)          # vision_transformer.py:137 out_indices


def trunc_normal_(t, std=0.02):
    # Dino/modules/utils.py:523-561 -- normal(0, std) truncated to [-2, 2] absolute (i.e. +-100 sigma at std=.02)
    with torch.no_grad():
        return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


def pos_resample_matrix():
    """interpolate_pos_encoding (vision_transformer.py:182-201) is linear in pos_embed: P_eff = W @ pos_embed.
    W [256 out tokens, 256 table entries] is a constant of the architecture; built once on the host by pushing the
    identity through the same bicubic scale-factor resample (16x16 -> 8x32, scale (8.1/16, 32.1/16))."""
    eye = torch.eye(256, dtype=torch.float64).reshape(1, 16, 16, 256).permute(0, 3, 1, 2)
    out = nn.functional.interpolate(eye, scale_factor=(8.1 / 16.0, 32.1 / 16.0), mode="bicubic")
    assert out.shape[-2:] == (8, 32)
    return out.permute(0, 2, 3, 1).reshape(256, 256).to(torch.float32)   # [token_out, table_in]


class _Attention(nn.Module):        # parameter container (names of vision_transformer.py:68-79)
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, drop_path):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.drop_prob = float(drop_path)


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


# per-block parameter order inside the flat argument list of VitFn
_BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
               "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")
_GEMM_W = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")


class VisionTransformer(nn.Module):
    """Drop-in for Dino/modules/vision_transformer.py:134 (img 32x128, patch 4, no CLS token in the sequence)."""

    def __init__(self, img_size=[32, 128], patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=None, out_indices=[2, 4, 6], **kwargs):
        super().__init__()
        if list(img_size) != [32, 128] or patch_size != 4 or in_chans != 3 or embed_dim % 64 or embed_dim // num_heads != 64 \
                or not qkv_bias or drop_rate or attn_drop_rate or list(out_indices) != [2, 4, 6] or embed_dim > 512:
            raise NotImplementedError("ccd_b200 implements the CCD configuration only: 32x128 crops, patch 4, head_dim 64, "
                                      "qkv_bias, E <= 512 (vit_tiny/small/base of the reference)")
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.out_indices = out_indices
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))          # unused by forward, kept for checkpoints
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]   # vision_transformer.py:150
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, dpr[i]) for i in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Identity()
        self.fc = nn.Identity()
        self.norm_seg = nn.Sequential(nn.LayerNorm(embed_dim, eps=1e-6), nn.LayerNorm(embed_dim, eps=1e-6),
                                      nn.LayerNorm(embed_dim, eps=1e-6))
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        for m in self.modules():                                             # _init_weights :174-180
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        # constant of the architecture, NOT a registered buffer: DistributedDataParallel re-broadcasts every buffer from rank 0
        # at each forward (256 KB here, on the critical path of the step)
        self._pos_w_host = pos_resample_matrix()
        self._pos_w_dev = {}
        self._cast = ops.ChunkTable()
        self._bf16 = None
        self._bf16_ver = None
        self._keep = None

    # ---- flat parameter list handed to the autograd.Function (fixed order) ----
    def _param_list(self):
        ps = [self.pos_embed, self.patch_embed.proj.weight, self.patch_embed.proj.bias]
        for blk in self.blocks:
            sd = dict(blk.named_parameters())
            ps += [sd[k] for k in _BLOCK_KEYS]
        ps += [self.norm.weight, self.norm.bias]
        for ln in self.norm_seg:
            ps += [ln.weight, ln.bias]
        return ps

    def _bf16_srcs(self):
        srcs = []
        for blk in self.blocks:
            sd = dict(blk.named_parameters())
            srcs += [sd[k] for k in _GEMM_W]
        return srcs

    def bf16_copies(self):
        """(parameter, bf16 operand copy) pairs, for the fused optimizer step that refreshes the copies in its own pass."""
        return list(zip(self._bf16_srcs(), self._bf16)) if self._bf16 is not None else []

    # ---- freshness of the bf16 operand copies -------------------------------------------------------------------------
    # A tensor's version counter does not see `.data` / raw-pointer / set_ updates (the reference's own teacher EMA is
    # `param_k.data.mul_(m).add_(...)`, train.py:264-272), so it cannot be the freshness signal.  Rule: EVERY forward
    # re-casts (one multi-tensor kernel, ~50 us for ViT-Small) unless the fused optimizer step (optim.AdamW), which
    # rewrites parameter and copy in the same pass, vouched for the copies since the last forward -- a one-shot token.
    def bf16_mark_fresh(self):
        self._bf16_ver = tuple(p._version for p in self._bf16_srcs())

    def bf16_invalidate(self):
        """Call after modifying GEMM weights behind the fused optimizer's back between its step() and the next forward."""
        self._bf16_ver = None

    def bf16_is_fresh(self):
        return self._bf16_ver is not None and self._bf16_ver == tuple(p._version for p in self._bf16_srcs())

    def _bf16_weights(self):
        """bf16 operand copies of the GEMM weights: one multi-tensor cast kernel per forward, skipped only when the fused
        optimizer step rewrote the copies itself (see above)."""
        params = self._bf16_srcs()
        srcs = [p.detach() for p in params]
        if self._bf16 is None or self._bf16[0].device != srcs[0].device:
            self._bf16 = [torch.empty(s.shape, dtype=torch.bfloat16, device=s.device) for s in srcs]
            self._bf16_ver = None
        if not self.bf16_is_fresh():
            table, n = self._cast.get(srcs, self._bf16, 2)
            ops.multi_tensor(ops.MT_CAST_BF16, table, n)
        self._bf16_ver = None                                  # one-shot: consumed by this forward
        E = self.embed_dim
        wp = torch.zeros(E, 64, dtype=torch.bfloat16, device=srcs[0].device)     # K padded 48 -> 64
        wp[:, :48] = self.patch_embed.proj.weight.detach().reshape(E, 48)
        return wp, self._bf16

    def pos_w(self, device):
        w = self._pos_w_dev.get(device)
        if w is None:
            w = self._pos_w_dev[device] = self._pos_w_host.to(device).contiguous()
        return w

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ccd_b200.VisionTransformer runs on CUDA (sm_100a) only; there is no CPU path")
        drop = None
        if self.training and any(b.drop_prob > 0 for b in self.blocks):
            n = x.shape[0]
            keep = self._keep
            if keep is None or keep.device != x.device:       # constant of the architecture: uploaded once (a pageable
                keep = self._keep = torch.tensor(             # H2D copy per forward would stall the host on the stream)
                    [1.0 - b.drop_prob for b in self.blocks for _ in range(2)], device=x.device).view(-1, 1)
            # drop_path (vision_transformer.py:27-36): floor(keep + U[0,1)) / keep per sample, separately per branch
            drop = (torch.floor(keep + torch.rand(2 * len(self.blocks), n, device=x.device)) / keep).contiguous()
        wp, wb = self._bf16_weights()
        outs = VitFn.apply(x.contiguous().float(), self, wp, wb, drop, *self._param_list())
        tokens, taps = outs[0], outs[1:]
        n = x.shape[0]
        E = self.embed_dim
        taps2d = [t.view(n, 8, 32, E).permute(0, 3, 1, 2) for t in taps]       # to_2D :237-238 (NHWC storage)
        return tokens.view(n, GRID_TOKENS, E), taps2d


class VitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_img, mod, wp, wb, drop, *params):
        E, H = mod.embed_dim, mod.num_heads
        n = x_img.shape[0]
        T = n * GRID_TOKENS
        dev = x_img.device
        save = any(ctx.needs_input_grad)     # grad mode is off inside Function.forward; this reflects the caller's
        f32 = dict(dtype=torch.float32, device=dev)
        b16 = dict(dtype=torch.bfloat16, device=dev)
        pos_embed, _, patch_b = params[0], params[1], params[2]
        cols = ops.patch_im2col(x_img)
        pos_eff = ops.smallmm(mod.pos_w(dev), pos_embed.detach()[0].contiguous())  # [256,E] = W . pos_embed, fp32
        x = torch.empty(T, E, **f32)
        ops.gemm(cols, wp, T, E, 64, 0, 0, ops.EPI_POS, patch_b.detach(), x, None, pos_eff)
        saved = []
        taps = []
        pi = 3
        for l, blk in enumerate(mod.blocks):
            (g1, b1, _, bqkv, _, bproj, g2, b2, _, bfc1, _, bfc2) = [p.detach() for p in params[pi:pi + 12]]
            pi += 12
            wqkv, wproj, wfc1, wfc2 = wb[4 * l: 4 * l + 4]
            ds_attn = drop[2 * l] if drop is not None else None
            ds_mlp = drop[2 * l + 1] if drop is not None else None
            xn, _ = ops.layernorm_fwd(x, g1, b1)
            qkv = torch.empty(T, 3 * E, **b16)
            ops.linear_fwd(xn, wqkv, bqkv, ops.EPI_BF16, qkv)
            o, lse = ops.mhsa_fwd(qkv, n, H, want_lse=save)
            x2 = torch.empty(T, E, **f32)
            ops.linear_fwd(o, wproj, bproj, ops.EPI_RESID, x2, None, x, seq_scale=ds_attn)
            xn2, _ = ops.layernorm_fwd(x2, g2, b2)
            hpre = torch.empty(T, 4 * E, **b16) if save else None      # pre-activation only feeds the backward
            hact = torch.empty(T, 4 * E, **b16)
            ops.linear_fwd(xn2, wfc1, bfc1, ops.EPI_GELU, hpre, hact)
            x3 = torch.empty(T, E, **f32)
            ops.linear_fwd(hact, wfc2, bfc2, ops.EPI_RESID, x3, None, x2, seq_scale=ds_mlp)
            if save:
                saved.append((x, xn, qkv, o, lse, x2, xn2, hpre, hact))
            x = x3
            if l + 1 in TAP_AFTER:
                i = TAP_AFTER.index(l + 1)
                gs, bs = params[3 + 12 * len(mod.blocks) + 2 + 2 * i].detach(), params[3 + 12 * len(mod.blocks) + 3 + 2 * i].detach()
                _, tap = ops.layernorm_fwd(x, gs, bs, want_bf16=False, want_f32=True)
                taps.append(tap)
                if save:
                    saved[-1] = saved[-1] + (x,)
        gn, bn = params[3 + 12 * len(mod.blocks)].detach(), params[3 + 12 * len(mod.blocks) + 1].detach()
        _, tokens = ops.layernorm_fwd(x, gn, bn, want_bf16=False, want_f32=True)
        if save:
            ctx.mod, ctx.wb, ctx.drop, ctx.cols, ctx.x_final, ctx.saved = mod, wb, drop, cols, x, saved
            ctx.params = params
            ctx.n = n
        return (tokens, *taps)

    @staticmethod
    def backward(ctx, d_tokens, *d_taps):
        mod, wb, drop, params = ctx.mod, ctx.wb, ctx.drop, ctx.params
        E, H, n = mod.embed_dim, mod.num_heads, ctx.n
        T = n * GRID_TOKENS
        dev = ctx.x_final.device
        depth = len(mod.blocks)
        b16 = dict(dtype=torch.bfloat16, device=dev)
        # one zero-filled flat buffer for every parameter gradient (split-K wgrads and LN/bias sums accumulate)
        sizes = [p.numel() for p in params]
        flat = torch.zeros(sum(sizes) + 64 * E, dtype=torch.float32, device=dev)
        grads, off = [], 0
        for p, s in zip(params, sizes):
            grads.append(flat[off:off + s].view(p.shape))
            off += s
        d_wpatch = flat[off:off + 64 * E].view(E, 64)

        def zero_like_tokens():
            return torch.zeros(T, E, dtype=torch.float32, device=dev)

        i_norm = 3 + 12 * depth
        d_tokens = d_tokens.contiguous().float().view(T, E) if d_tokens is not None else zero_like_tokens()
        last_scale = drop[2 * depth - 1] if drop is not None else None
        # every LN backward also accumulates the column sums of its bf16 output = bias gradient of the layer fed by it
        dx, dxb = ops.layernorm_bwd(ctx.x_final, params[i_norm].detach(), d_tokens, None, grads[i_norm], grads[i_norm + 1],
                                    bf16_seq_scale=last_scale, dbias_next=grads[3 + 12 * (depth - 1) + 11])
        for l in reversed(range(depth)):
            sv = ctx.saved[l]
            x, xn, qkv, o, lse, x2, xn2, hpre, hact = sv[:9]
            pi = 3 + 12 * l
            g1, g2 = params[pi].detach(), params[pi + 6].detach()
            wqkv, wproj, wfc1, wfc2 = wb[4 * l: 4 * l + 4]
            ds_attn = drop[2 * l] if drop is not None else None
            ds_mlp = drop[2 * l + 1] if drop is not None else None
            if l + 1 in TAP_AFTER:
                i = TAP_AFTER.index(l + 1)
                dt = d_taps[i]
                if dt is not None:
                    dt = dt.contiguous().float().view(T, E)
                    gi = i_norm + 2 + 2 * i
                    grads[pi + 11].zero_()        # fc2.bias sums restart: this tap re-derives the bf16 copy
                    dx, dxb = ops.layernorm_bwd(sv[9], params[gi].detach(), dt, dx, grads[gi], grads[gi + 1], bf16_seq_scale=ds_mlp,
                                                dbias_next=grads[pi + 11])
            # ---- MLP branch: x3 = x2 + s*(gelu(xn2 W1^T + b1) W2^T + b2);  dxb = s * dx (bf16) ----
            ops.linear_wgrad(dxb, hact, grads[pi + 10])             # fc2.bias grad came from the producer of dxb
            dh = torch.empty(T, 4 * E, **b16)
            ops.linear_dgrad(dxb, wfc2, ops.EPI_DGELU, dh, hpre, colsum=grads[pi + 9])    # fc1.bias grad in the epilogue
            ops.linear_wgrad(dh, xn2, grads[pi + 8])
            dxn2 = torch.empty(T, E, **b16)
            ops.linear_dgrad(dh, wfc1, ops.EPI_BF16, dxn2)
            dx2, dx2b = ops.layernorm_bwd(x2, g2, dxn2, dx, grads[pi + 6], grads[pi + 7], bf16_seq_scale=ds_attn,
                                          dbias_next=grads[pi + 5])
            # ---- attention branch: x2 = x + s*(attn(xn) Wp^T + bp) ----
            ops.linear_wgrad(dx2b, o, grads[pi + 4])                # proj.bias grad came from the LN2 backward above
            d_o = torch.empty(T, E, **b16)
            ops.linear_dgrad(dx2b, wproj, ops.EPI_BF16, d_o)
            # qkv.bias grad: q part inside the kernel, v part = (proj.bias grad) @ W_proj (column sums of d_o), k part = 0
            dqkv = ops.mhsa_bwd(qkv, o, d_o, lse, n, H, dbias=grads[pi + 3], dproj_bias=grads[pi + 5],
                                w_proj=params[pi + 4].detach())
            ops.linear_wgrad(dqkv, xn, grads[pi + 2])
            dxn = torch.empty(T, E, **b16)
            ops.linear_dgrad(dqkv, wqkv, ops.EPI_BF16, dxn)
            prev_scale = None
            if drop is not None and l > 0:
                prev_scale = drop[2 * l - 1]
            # NB: a tap LN backward (if any) sits between this block and the previous one and re-derives the bf16 copy
            dx, dxb = ops.layernorm_bwd(x, g1, dxn, dx2, grads[pi], grads[pi + 1], bf16_seq_scale=prev_scale,
                                        dbias_next=(grads[pi - 12 + 11] if l > 0 else grads[2]))
            ctx.saved[l] = None
        # ---- patch embed: x0 = cols Wp^T + b + P_eff ----
        ops.linear_wgrad(dxb, ctx.cols, d_wpatch)                   # patch bias grad came from block 0's LN1 backward
        grads[1].copy_(d_wpatch[:, :48].reshape(grads[1].shape))
        dpos_eff = torch.zeros(GRID_TOKENS * E, dtype=torch.float32, device=dev)
        ops.colsum_f32(dx.view(n, GRID_TOKENS * E), dpos_eff)
        grads[0].copy_(ops.smallmm(mod.pos_w(dev), dpos_eff.view(GRID_TOKENS, E), trans_a=True).view(grads[0].shape))
        return (None, None, None, None, None, *grads)


def vit_tiny(patch_size=16, **kwargs):          # vision_transformer.py:273-277
    return VisionTransformer(patch_size=patch_size, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, **kwargs)


def vit_small(patch_size=16, **kwargs):         # :280-284
    return VisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True, **kwargs)


def vit_base(patch_size=16, **kwargs):          # :287-291  (E=512, 8 heads in this repository)
    return VisionTransformer(patch_size=patch_size, embed_dim=512, depth=12, num_heads=8, mlp_ratio=4, qkv_bias=True, **kwargs)
