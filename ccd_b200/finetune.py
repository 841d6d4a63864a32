"""Recognition / fine-tuning path of CCD on the sm_100a kernels (BASELINE config 5).

Drop-in for `DINO_Finetune` (Dino/model/dino_vision.py:135-290), `NRTRDecoder` (Dino/decoder/nrtr_decoder.py:12-203),
`TFDecoderLayer` / `MultiHeadAttention` / `PositionwiseFeedForward` / `PositionalEncoding`
(Dino/decoder/transformer_layers.py:73-163, transformer_module.py:37-162), `TFLoss` (Dino/loss/ce_loss.py:94-128) and
`AttnConvertor` (Dino/convertor/attn.py:7-141): same constructors, attribute and parameter names (reference checkpoints
load), same forward signatures and outputs -- but every contraction is a tcgen05 GEMM of libccd_b200.so with a fused
epilogue (GELU, residual add, GELU'), the decoder's attentions are one fused kernel per (sample, head) whose probabilities
never reach HBM, the 12 cross-attention K/V projections of the 6 layers are ONE GEMM over the encoder memory, and the
greedy `forward_test` projects the encoder memory once and decodes incrementally with a per-layer key / value cache
instead of re-running the whole decoder at every step.

  img [N,3,32,128] -> VisionTransformer (ccd_b200.encoder) -> Mlp E->512->512 -> memory bf16 [N*256, 512]
  targets [N,T] -> embedding + sinusoid table -> 6 x (LN -> masked self-attention -> +res -> LN -> cross-attention -> +res
                   -> LN -> FFN 512->256->512 -> +res) -> LN -> Linear 512->92 -> TFLoss
There is no CPU / eager fallback: the modules raise on CPU tensors.
"""

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .encoder import VisionTransformer, vit_base, vit_small, vit_tiny  # noqa: F401

N_HEAD, D_HEAD = 8, 64


def _zeros(*shape, like, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------------------
# autograd building blocks over the C ABI
# ------------------------------------------------------------------------------------------------------------
class CastBf16Fn(torch.autograd.Function):
    """fp32 activations -> bf16 GEMM operand."""

    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


class DropoutFn(torch.autograd.Function):
    """nn.Dropout with a regenerable mask (ccd_dropout): the backward replays the same (seed, index) mask on dy."""

    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout(x.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy.contiguous(), ctx.p, ctx.seed), None, None


class MultiLinearBf16Fn(torch.autograd.Function):
    """y bf16 [T, sum N_i] = x bf16 [T,K] @ cat(W_i)^T (+ bias): several bias-free nn.Linear layers that share their input
    (linear_q / linear_k / linear_v, transformer_module.py:62-64,76-78) as ONE GEMM over the concatenated bf16 weight."""

    @staticmethod
    def forward(ctx, x, w_cat_b16, bias, *weights):
        T, K = x.shape
        N = w_cat_b16.shape[0]
        y = torch.empty(T, N, dtype=torch.bfloat16, device=x.device)
        ops.gemm(x, w_cat_b16, T, N, K, 0, 0, ops.EPI_BF16, bias.detach() if bias is not None else None, y)
        ctx.save_for_backward(x, w_cat_b16)
        ctx.sizes = [w.shape[0] for w in weights]
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        T, K = x.shape
        N = w.shape[0]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(T, K, dtype=torch.bfloat16, device=x.device)
            ops.linear_dgrad(dy, w, ops.EPI_BF16, dx)
        dw = _zeros(N, K, like=x)
        ops.linear_wgrad(dy, x, dw)
        db = ops.colsum_bf16(dy, _zeros(N, like=x)) if ctx.has_bias else None
        return (dx, None, db) + tuple(dw.split(ctx.sizes, dim=0))


class LinearResidFn(torch.autograd.Function):
    """x_new f32 = x_res + dropout(o bf16 @ W^T + b): MultiHeadAttention.fc + proj_drop + the residual add of the layer
    (transformer_module.py:92-93, transformer_layers.py:150-156).  p = 0: one GEMM with the residual epilogue."""

    @staticmethod
    def forward(ctx, o, w_b16, weight, bias, x_res, p, seed):
        T, K = o.shape
        N = w_b16.shape[0]
        out = torch.empty(T, N, dtype=torch.float32, device=o.device)
        b = bias.detach() if bias is not None else None
        if p > 0:
            y = torch.empty(T, N, dtype=torch.float32, device=o.device)
            ops.gemm(o, w_b16, T, N, K, 0, 0, ops.EPI_F32, b, y)
            out = ops.dropout(y, p, seed, resid=x_res.contiguous())
        else:
            ops.gemm(o, w_b16, T, N, K, 0, 0, ops.EPI_RESID, b, out, None, x_res.contiguous())
        ctx.save_for_backward(o, w_b16)
        ctx.p, ctx.seed, ctx.has_bias = p, seed, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        o, w = ctx.saved_tensors
        g = g.contiguous()
        gd = ops.dropout(g, ctx.p, ctx.seed) if ctx.p > 0 else g
        gb = ops.cast_bf16(gd)
        T, K = o.shape
        N = w.shape[0]
        do = torch.empty(T, K, dtype=torch.bfloat16, device=o.device)
        ops.linear_dgrad(gb, w, ops.EPI_BF16, do)
        dw = _zeros(N, K, like=o)
        ops.linear_wgrad(gb, o, dw)
        db = ops.colsum_f32(gd, _zeros(N, like=o)) if ctx.has_bias else None
        return do, None, dw, db, g, None, None


class FFNFn(torch.autograd.Function):
    """out = [x_res +] dropout_out( W2 dropout_mid(gelu(W1 x + b1)) + b2 ).
    PositionwiseFeedForward (transformer_module.py:99-128: p_mid = 0, residual, f32 out) and DINO_Finetune.encoder = Mlp
    (dino_vision.py:117-133: p_mid = p_out = 0.1, no residual, bf16 out).  GELU in the fc1 epilogue, GELU' in the fc2-dgrad
    epilogue, residual add in the fc2 epilogue."""

    @staticmethod
    def forward(ctx, x, w1_b16, w2_b16, w1, b1, w2, b2, x_res, p_mid, p_out, seed):
        T, K = x.shape
        Hd = w1_b16.shape[0]
        N = w2_b16.shape[0]
        dev = x.device
        pre = torch.empty(T, Hd, dtype=torch.bfloat16, device=dev)
        act = torch.empty(T, Hd, dtype=torch.bfloat16, device=dev)
        ops.gemm(x, w1_b16, T, Hd, K, 0, 0, ops.EPI_GELU, b1.detach(), pre, act)
        act_d = ops.dropout(act, p_mid, seed) if p_mid > 0 else act
        if x_res is not None:
            if p_out > 0:
                y = torch.empty(T, N, dtype=torch.float32, device=dev)
                ops.gemm(act_d, w2_b16, T, N, Hd, 0, 0, ops.EPI_F32, b2.detach(), y)
                out = ops.dropout(y, p_out, seed + 1, resid=x_res.contiguous())
            else:
                out = torch.empty(T, N, dtype=torch.float32, device=dev)
                ops.gemm(act_d, w2_b16, T, N, Hd, 0, 0, ops.EPI_RESID, b2.detach(), out, None, x_res.contiguous())
        else:
            out = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
            ops.gemm(act_d, w2_b16, T, N, Hd, 0, 0, ops.EPI_BF16, b2.detach(), out)
            if p_out > 0:
                out = ops.dropout(out, p_out, seed + 1)
        ctx.save_for_backward(x, w1_b16, w2_b16, pre, act_d)
        ctx.cfg = (p_mid, p_out, seed, x_res is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w1, w2, pre, act_d = ctx.saved_tensors
        p_mid, p_out, seed, has_res = ctx.cfg
        g = g.contiguous()
        T, K = x.shape
        Hd, N = w1.shape[0], w2.shape[0]
        gd = ops.dropout(g, p_out, seed + 1) if p_out > 0 else g
        if gd.dtype == torch.float32:
            db2 = ops.colsum_f32(gd, _zeros(N, like=x))
            gb = ops.cast_bf16(gd)
        else:
            gb = gd
            db2 = ops.colsum_bf16(gb, _zeros(N, like=x))
        dw2 = _zeros(N, Hd, like=x)
        ops.linear_wgrad(gb, act_d, dw2)
        dpre = torch.empty(T, Hd, dtype=torch.bfloat16, device=x.device)
        db1 = _zeros(Hd, like=x)
        if p_mid > 0:
            ops.linear_dgrad(gb, w2, ops.EPI_DGELU, dpre, aux=pre)            # acc * gelu'(pre), then the activation-dropout mask
            dpre = ops.dropout(dpre, p_mid, seed)
            ops.colsum_bf16(dpre, db1)
        else:
            ops.linear_dgrad(gb, w2, ops.EPI_DGELU, dpre, aux=pre, colsum=db1)
        dw1 = _zeros(Hd, K, like=x)
        ops.linear_wgrad(dpre, x, dw1)
        dx = torch.empty(T, K, dtype=torch.bfloat16, device=x.device)
        ops.linear_dgrad(dpre, w1, ops.EPI_BF16, dx)
        return dx, None, None, dw1, db1, dw2, db2, (g if has_res else None), None, None, None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dimension (<= 512) of the fp32 residual stream -> bf16 GEMM operand."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        y, _ = ops.layernorm_fwd(x, gamma.detach(), beta.detach(), eps=eps)
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        E = x.shape[1]
        dg, db = _zeros(E, like=x), _zeros(E, like=x)
        dx, _ = ops.layernorm_bwd(x, gamma.detach(), dy.contiguous(), None, dg, db, want_f32=True, want_bf16=False, eps=ctx.eps)
        return dx, dg, db, None


class DecSelfAttnFn(torch.autograd.Function):
    """Masked self-attention over the target tokens from the fused projection qkv bf16 [N*T, 3*512] (q | k | v)."""

    @staticmethod
    def forward(ctx, qkv, trg, n, t, pad_idx, p, seed):
        D = N_HEAD * D_HEAD
        o, lse = ops.dec_attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], n, N_HEAD, t, t, trg, pad_idx, p, seed)
        ctx.save_for_backward(qkv, o, lse, trg)
        ctx.cfg = (n, t, pad_idx, p, seed)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse, trg = ctx.saved_tensors
        n, t, pad_idx, p, seed = ctx.cfg
        D = N_HEAD * D_HEAD
        dqkv = torch.empty_like(qkv)
        ops.dec_attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do.contiguous(), lse, dqkv[:, :D], dqkv[:, D:2 * D],
                         dqkv[:, 2 * D:], n, N_HEAD, t, t, trg, pad_idx, p, seed)
        return dqkv, None, None, None, None, None, None


class DecCrossAttnFn(torch.autograd.Function):
    """Cross-attention of the T target positions over the 256 encoder tokens; kv = (k | v) bf16 column view [N*256, 2*512]."""

    @staticmethod
    def forward(ctx, q, kv, n, t, tk, p, seed):
        D = N_HEAD * D_HEAD
        o, lse = ops.dec_attn_fwd(q, kv[:, :D], kv[:, D:], n, N_HEAD, t, tk, None, 0, p, seed)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.cfg = (n, t, tk, p, seed)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        n, t, tk, p, seed = ctx.cfg
        D = N_HEAD * D_HEAD
        dq = torch.empty_like(q)
        dkv = torch.empty(kv.shape, dtype=torch.bfloat16, device=kv.device)
        ops.dec_attn_bwd(q, kv[:, :D], kv[:, D:], o, do.contiguous(), lse, dq, dkv[:, :D], dkv[:, D:], n, N_HEAD, t, tk, None, 0, p, seed)
        return dq, dkv, None, None, None, None, None


class ClassifierFn(torch.autograd.Function):
    """logits f32 [T, 96] = hid bf16 [T,512] @ Wpad^T + bpad: NRTRDecoder.classifier (nrtr_decoder.py:79-80,150) with the 92
    classes padded to a multiple of 8 columns (zero weight rows; the loss reads the first 92)."""

    @staticmethod
    def forward(ctx, hid, w_pad_b16, b_pad, weight, bias):
        T, K = hid.shape
        NP = w_pad_b16.shape[0]
        y = torch.empty(T, NP, dtype=torch.float32, device=hid.device)
        ops.gemm(hid, w_pad_b16, T, NP, K, 0, 0, ops.EPI_F32, b_pad, y)
        ctx.save_for_backward(hid, w_pad_b16)
        ctx.n_cls = weight.shape[0]
        return y

    @staticmethod
    def backward(ctx, g):
        hid, w = ctx.saved_tensors
        g = g.contiguous()
        T, K = hid.shape
        NP = w.shape[0]
        gb = ops.cast_bf16(g)
        dh = torch.empty(T, K, dtype=torch.bfloat16, device=hid.device)
        ops.linear_dgrad(gb, w, ops.EPI_BF16, dh)
        dw = _zeros(NP, K, like=hid)
        ops.linear_wgrad(gb, hid, dw)
        db = ops.colsum_f32(g, _zeros(NP, like=hid))
        return dh, None, None, dw[:ctx.n_cls], db[:ctx.n_cls]


class TFLossFn(torch.autograd.Function):
    """TFLoss.forward (ce_loss.py:94-128 + CELoss.forward :41-59): mean cross-entropy of logits[:, :-1] against
    targets[:, 1:] with PAD ignored; one kernel produces the loss sum, the counted rows and softmax - onehot."""

    @staticmethod
    def forward(ctx, logits, targets, n_classes, pad_idx):
        acc, dl = ops.tf_ce(logits, n_classes, targets, pad_idx)
        ctx.save_for_backward(dl, acc)
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        dl, acc = ctx.saved_tensors
        return dl * (g / acc[1]), None, None, None


# ------------------------------------------------------------------------------------------------------------
# parameter containers with the reference's names
# ------------------------------------------------------------------------------------------------------------
class Mlp(nn.Module):                                   # dino_vision.py:117-133
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class MultiHeadAttention(nn.Module):                    # transformer_module.py:37-96
    def __init__(self, n_head=8, d_model=512, d_k=64, d_v=64, dropout=0.1, qkv_bias=False):
        super().__init__()
        if n_head != N_HEAD or d_k != D_HEAD or d_v != D_HEAD or qkv_bias or d_model != n_head * d_k:
            raise NotImplementedError("ccd_b200 implements the CCD decoder configuration: 8 heads of 64, d_model 512, no qkv bias")
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.dim_k = self.dim_v = n_head * d_k
        self.linear_q = nn.Linear(self.dim_k, self.dim_k, bias=False)
        self.linear_k = nn.Linear(self.dim_k, self.dim_k, bias=False)
        self.linear_v = nn.Linear(self.dim_v, self.dim_v, bias=False)
        self.fc = nn.Linear(self.dim_v, d_model, bias=False)
        self.dropout_p = dropout                        # ScaledDotProductAttention.dropout and proj_drop


class PositionwiseFeedForward(nn.Module):               # transformer_module.py:99-128
    def __init__(self, d_in, d_hid, dropout=0.1, act_cfg=None):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.act = nn.GELU()
        self.dropout_p = dropout


class PositionalEncoding(nn.Module):                    # transformer_module.py:131-162
    def __init__(self, d_hid=512, n_position=200, dropout=0):
        super().__init__()
        self.register_buffer("position_table", self._get_sinusoid_encoding_table(n_position, d_hid))

    @staticmethod
    def _get_sinusoid_encoding_table(n_position, d_hid):
        denominator = torch.Tensor([1.0 / np.power(10000, 2 * (hid_j // 2) / d_hid) for hid_j in range(d_hid)]).view(1, -1)
        table = torch.arange(n_position).unsqueeze(-1).float() * denominator
        table[:, 0::2] = torch.sin(table[:, 0::2])
        table[:, 1::2] = torch.cos(table[:, 1::2])
        return table.unsqueeze(0)


class TFDecoderLayer(nn.Module):                        # transformer_layers.py:73-163 (default pre-LN operation order)
    def __init__(self, d_model=512, d_inner=256, n_head=8, d_k=64, d_v=64, dropout=0.1, qkv_bias=False, act_cfg=None,
                 operation_order=None):
        super().__init__()
        if operation_order not in (None, ('norm', 'self_attn', 'norm', 'enc_dec_attn', 'norm', 'ffn')):
            raise NotImplementedError("ccd_b200 implements the pre-LN operation order the reference uses")
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.self_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, qkv_bias=qkv_bias)
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, qkv_bias=qkv_bias)
        self.mlp = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)


class NRTRDecoder(nn.Module):
    """Drop-in for Dino/decoder/nrtr_decoder.py:12 (training forward + greedy decoding) on the C-ABI kernels."""

    def __init__(self, n_layers=6, d_embedding=512, n_head=8, d_k=64, d_v=64, d_model=512, d_inner=256, n_position=200,
                 dropout=0.1, num_classes=93, max_seq_len=40, start_idx=1, padding_idx=92, init_cfg=None, **kwargs):
        super().__init__()
        if d_embedding != d_model or d_model != 512 or max_seq_len + 1 > 32 or d_inner % 64:
            raise NotImplementedError("ccd_b200 decoder: d_model = d_embedding = 512, max_seq_len <= 31, d_inner % 64 == 0")
        self.padding_idx, self.start_idx, self.max_seq_len = padding_idx, start_idx, max_seq_len
        self.d_model, self.dropout_p = d_model, dropout
        self.trg_word_emb = nn.Embedding(num_classes, d_embedding, padding_idx=padding_idx)
        self.position_enc = PositionalEncoding(d_embedding, n_position=n_position)
        self.layer_stack = nn.ModuleList([TFDecoderLayer(d_model, d_inner, n_head, d_k, d_v, dropout=dropout, **kwargs)
                                          for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.classifier = nn.Linear(d_model, num_classes - 1)          # PAD is never predicted
        self._cast = ops.ChunkTable()
        self._buf = None
        self._calls = 0

    # ---- bf16 operand copies: concatenated where layers share their input; one multi-tensor cast per step ----
    def _groups(self):
        g = {"kv_all": [w for l in self.layer_stack for w in (l.enc_attn.linear_k.weight, l.enc_attn.linear_v.weight)]}
        for i, l in enumerate(self.layer_stack):
            g[f"qkv{i}"] = [l.self_attn.linear_q.weight, l.self_attn.linear_k.weight, l.self_attn.linear_v.weight]
            g[f"sfc{i}"] = [l.self_attn.fc.weight]
            g[f"cq{i}"] = [l.enc_attn.linear_q.weight]
            g[f"cfc{i}"] = [l.enc_attn.fc.weight]
            g[f"w1{i}"] = [l.mlp.w_1.weight]
            g[f"w2{i}"] = [l.mlp.w_2.weight]
        g["cls"] = [self.classifier.weight]
        return g

    def bf16_weights(self):
        groups = self._groups()
        params = [p for ws in groups.values() for p in ws]
        dev = params[0].device
        if self._buf is None or next(iter(self._buf.values())).device != dev:
            self._buf = {}
            for name, ws in groups.items():
                rows = sum(w.shape[0] for w in ws)
                if name == "cls":
                    rows = (rows + 7) // 8 * 8
                self._buf[name] = torch.zeros(rows, ws[0].shape[1], dtype=torch.bfloat16, device=dev)
        # re-cast on every forward (one small multi-tensor launch): tensor version counters do not see `.data` / raw-pointer
        # updates, so they cannot vouch for the copies
        srcs, dsts = [], []
        for name, ws in groups.items():
            r = 0
            for w in ws:
                srcs.append(w.detach())
                dsts.append(self._buf[name][r:r + w.shape[0]])
                r += w.shape[0]
        table, n = self._cast.get(srcs, dsts, 2)
        ops.multi_tensor(ops.MT_CAST_BF16, table, n)
        return self._buf

    def _seed(self):
        """Seed of one dropout site: a fresh draw from torch's CPU generator (so torch.manual_seed / RNG-state restores are
        honoured and a resumed run does not replay masks) mixed with the rank (DDP ranks hold the same torch seed but must
        drop independently, as the reference's per-rank CUDA generators do)."""
        self._calls += 1
        draw = int(torch.randint(0, 1 << 62, (1,)).item())
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        return (draw ^ (rank * 0x9E3779B97F4A7C15) ^ (self._calls * 7919)) & 0x7FFFFFFFFFFFFFFF

    def project_memory(self, mem, wb):
        """K / V projections of ALL layers' cross-attentions over the encoder memory bf16 [N*256, 512]: one GEMM, N = 12*512."""
        ws = [w for l in self.layer_stack for w in (l.enc_attn.linear_k.weight, l.enc_attn.linear_v.weight)]
        return MultiLinearBf16Fn.apply(mem, wb["kv_all"], None, *ws)

    def hidden(self, trg_seq, kv_all, n_src_tokens, wb):
        """NRTRDecoder._attention (nrtr_decoder.py:98-116): returns the final-LN hidden states bf16 [N*T, 512]."""
        n, t = trg_seq.shape
        p = self.dropout_p if self.training else 0.0
        D = self.d_model
        emb = F.embedding(trg_seq, self.trg_word_emb.weight, self.padding_idx)                       # :99
        x = (emb + self.position_enc.position_table[:, :t].detach()).reshape(n * t, D).contiguous()  # :100
        if p > 0:
            x = DropoutFn.apply(x, p, self._seed())                                                  # :101
        kvs = kv_all.split(2 * D, dim=1)
        for i, l in enumerate(self.layer_stack):
            xn = LayerNormFn.apply(x, l.norm1.weight, l.norm1.bias, l.norm1.eps)
            sa = l.self_attn
            qkv = MultiLinearBf16Fn.apply(xn, wb[f"qkv{i}"], None, sa.linear_q.weight, sa.linear_k.weight, sa.linear_v.weight)
            o = DecSelfAttnFn.apply(qkv, trg_seq, n, t, self.padding_idx, p, self._seed())
            x = LinearResidFn.apply(o, wb[f"sfc{i}"], sa.fc.weight, None, x, p, self._seed())
            xn = LayerNormFn.apply(x, l.norm2.weight, l.norm2.bias, l.norm2.eps)
            ca = l.enc_attn
            q = MultiLinearBf16Fn.apply(xn, wb[f"cq{i}"], None, ca.linear_q.weight)
            o = DecCrossAttnFn.apply(q, kvs[i], n, t, n_src_tokens, p, self._seed())
            x = LinearResidFn.apply(o, wb[f"cfc{i}"], ca.fc.weight, None, x, p, self._seed())
            xn = LayerNormFn.apply(x, l.norm3.weight, l.norm3.bias, l.norm3.eps)
            m = l.mlp
            x = FFNFn.apply(xn, wb[f"w1{i}"], wb[f"w2{i}"], m.w_1.weight, m.w_1.bias, m.w_2.weight, m.w_2.bias, x, 0.0, p,
                            self._seed())
        return LayerNormFn.apply(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)

    def classify(self, hid, wb):
        npad = wb["cls"].shape[0]
        b_pad = torch.zeros(npad, dtype=torch.float32, device=hid.device)
        b_pad[: self.classifier.bias.shape[0]] = self.classifier.bias.detach()
        return ClassifierFn.apply(hid, wb["cls"], b_pad, self.classifier.weight, self.classifier.bias)

    def forward_train(self, feat, out_enc, targets_dict, img_metas=None):
        """out_enc = encoder memory bf16 [N*256, 512] (or [N,256,512]); returns (logits f32 [N,T,92], None)."""
        targets = targets_dict['padded_targets'].to(out_enc.device)
        n, t = targets.shape
        mem = out_enc.reshape(-1, self.d_model)
        wb = self.bf16_weights()
        kv_all = self.project_memory(mem, wb)
        logits = self.classify(self.hidden(targets, kv_all, mem.shape[0] // n, wb), wb)
        return logits.view(n, t, -1)[:, :, : self.classifier.weight.shape[0]], None

    @torch.no_grad()
    def forward_test(self, feat, out_enc, img_metas=None, test_speed=False, kv_cache=True):
        """Greedy decoding (nrtr_decoder.py:154-203): per-step softmax [N, max_seq_len, 92].
        The reference re-runs the whole decoder on the growing (PAD-filled) sequence at every step -- O(T^2) decoder passes of
        which only row `step` is used.  Causal + pad masking makes the hidden states of earlier positions independent of later
        tokens, so here (kv_cache=True) every step pushes ONE new token per sample through the layers, appending its
        self-attention key / value to a per-layer cache; the cross-attention K / V of all layers are projected once.
        kv_cache=False keeps the reference's schedule (used by the parity test of the cache)."""
        mem = out_enc.reshape(-1, self.d_model)
        n = mem.shape[0] // 256
        wb = self.bf16_weights()
        kv_all = self.project_memory(mem, wb)
        T1 = self.max_seq_len + 1
        seq = torch.full((n, T1), self.padding_idx, device=mem.device, dtype=torch.long)
        seq[:, 0] = self.start_idx
        n_cls = self.classifier.weight.shape[0]
        outputs = []
        was_training = self.training
        self.eval()
        D = self.d_model
        if kv_cache:
            kvs = kv_all.split(2 * D, dim=1)
            cache = [torch.zeros(n * T1, 3 * D, dtype=torch.bfloat16, device=mem.device) for _ in self.layer_stack]
            pos = self.position_enc.position_table[0]
        for step in range(self.max_seq_len):
            if kv_cache:
                tok = seq[:, step]
                x = (F.embedding(tok, self.trg_word_emb.weight, self.padding_idx) + pos[step]).contiguous()      # [N, 512] f32
                for i, l in enumerate(self.layer_stack):
                    sa, ca, m = l.self_attn, l.enc_attn, l.mlp
                    xn = LayerNormFn.apply(x, l.norm1.weight, l.norm1.bias, l.norm1.eps)
                    qkv = MultiLinearBf16Fn.apply(xn, wb[f"qkv{i}"], None, sa.linear_q.weight, sa.linear_k.weight, sa.linear_v.weight)
                    c = cache[i].view(n, T1, 3 * D)
                    c[:, step] = qkv                                                        # keys / values of positions <= step
                    keys = c[:, : step + 1].reshape(n * (step + 1), 3 * D)                  # every cached token is a real one (PAD is never predicted)
                    o, _ = ops.dec_attn_fwd(qkv[:, :D], keys[:, D:2 * D], keys[:, 2 * D:], n, N_HEAD, 1, step + 1, None, 0, 0.0, 0,
                                            want_lse=False)
                    x = LinearResidFn.apply(o, wb[f"sfc{i}"], sa.fc.weight, None, x, 0.0, 0)
                    xn = LayerNormFn.apply(x, l.norm2.weight, l.norm2.bias, l.norm2.eps)
                    q = MultiLinearBf16Fn.apply(xn, wb[f"cq{i}"], None, ca.linear_q.weight)
                    o, _ = ops.dec_attn_fwd(q, kvs[i][:, :D], kvs[i][:, D:], n, N_HEAD, 1, 256, None, 0, 0.0, 0, want_lse=False)
                    x = LinearResidFn.apply(o, wb[f"cfc{i}"], ca.fc.weight, None, x, 0.0, 0)
                    xn = LayerNormFn.apply(x, l.norm3.weight, l.norm3.bias, l.norm3.eps)
                    x = FFNFn.apply(xn, wb[f"w1{i}"], wb[f"w2{i}"], m.w_1.weight, m.w_1.bias, m.w_2.weight, m.w_2.bias, x, 0.0, 0.0, 0)
                hid = LayerNormFn.apply(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
                logits = self.classify(hid, wb)[:, :n_cls]
            else:
                hid = self.hidden(seq, kv_all, 256, wb)
                logits = self.classify(hid, wb).view(n, T1, -1)[:, step, :n_cls]
            prob = torch.softmax(logits, dim=-1)
            outputs.append(prob)
            seq[:, step + 1] = prob.argmax(dim=-1)
            if test_speed and int(prob.argmax()) == 91:                                        # :199-200, as written there
                break
        self.train(was_training)
        return torch.stack(outputs, dim=1)

    def forward(self, feat, out_enc, targets_dict=None, img_metas=None, train_mode=True, test_speed=False):   # base_decoder.py:19-31
        if train_mode:
            return self.forward_train(feat, out_enc, targets_dict, img_metas)
        return self.forward_test(feat, out_enc, img_metas, test_speed=test_speed)


class TFLoss(nn.Module):
    """Drop-in for Dino/loss/ce_loss.py:94 (flatten=True, reduction='mean')."""

    def __init__(self, ignore_index=-1, reduction='mean', flatten=True, **kwargs):
        super().__init__()
        if reduction != 'mean' or not flatten:
            raise NotImplementedError("ccd_b200.TFLoss implements the configuration DINO_Finetune uses (mean, flatten)")
        self.ignore_index = ignore_index

    def forward(self, outputs, targets_dict, img_metas=None):
        targets = targets_dict['padded_targets'].to(outputs.device)
        n, t, c = outputs.shape
        base = outputs._base if outputs._base is not None else outputs          # the padded [N*T, 96] logits of the decoder
        if (base.dim() == 2 and base.shape[0] == n * t and base.is_contiguous() and base.dtype == torch.float32
                and outputs.storage_offset() == base.storage_offset()
                and outputs.stride() == (t * base.shape[1], base.shape[1], 1)):
            flat = base
        else:                                                                   # foreign logits: pad the class dimension to 8
            flat = F.pad(outputs.reshape(n * t, c), (0, (-c) % 8)).contiguous()
        return TFLossFn.apply(flat.float(), targets.contiguous(), c, self.ignore_index)


class AttnConvertor:
    """Drop-in for Dino/convertor/attn.py:7-141 + base.py (host-side label <-> index conversion; plain Python like there)."""
    dicts = dict(
        DICT36=tuple('0123456789abcdefghijklmnopqrstuvwxyz'),
        DICT90=tuple('0123456789abcdefghijklmnopqrstuvwxyz' 'ABCDEFGHIJKLMNOPQRSTUVWXYZ!"#$%&\'()' '*+,-./:;<=>?@[\\]_`~'),
        DICT37=tuple('0123456789abcdefghijklmnopqrstuvwxyz '),
        DICT91=tuple('0123456789abcdefghijklmnopqrstuvwxyz' 'ABCDEFGHIJKLMNOPQRSTUVWXYZ!"#$%&\'()' '*+,-./:;<=>?@[\\]_`~ '))

    def __init__(self, dict_type='DICT90', dict_file=None, dict_list=None, with_unknown=True, max_seq_len=40, lower=False,
                 start_end_same=True, **kwargs):
        if dict_file is not None:
            with open(dict_file, encoding='utf-8') as f:
                self.idx2char = [ln.strip('\r\n') for ln in f if ln.strip('\r\n') != '']
        elif dict_list is not None:
            self.idx2char = list(dict_list)
        else:
            self.idx2char = list(self.dicts[dict_type])
        assert len(set(self.idx2char)) == len(self.idx2char), 'Invalid dictionary: Has duplicated characters.'
        self.with_unknown, self.max_seq_len, self.lower, self.start_end_same = with_unknown, max_seq_len, lower, start_end_same
        self.unknown_idx = None
        if with_unknown:
            self.idx2char.append('<UKN>')
            self.unknown_idx = len(self.idx2char) - 1
        self.idx2char.append('<BOS/EOS>')
        self.start_idx = len(self.idx2char) - 1
        if not start_end_same:
            self.idx2char.append('<BOS/EOS>')
        self.end_idx = len(self.idx2char) - 1
        self.idx2char.append('<PAD>')
        self.padding_idx = len(self.idx2char) - 1
        self.char2idx = {c: i for i, c in enumerate(self.idx2char)}

    def num_classes(self):
        return len(self.idx2char)

    def str2idx(self, strings):
        out = []
        for s in strings:
            if self.lower:
                s = s.lower()
            idx = []
            for ch in s:
                ci = self.char2idx.get(ch, self.unknown_idx)
                if ci is None:
                    raise Exception(f'Chararcter: {ch} not in dict, please check gt_label and use custom dict file, or set '
                                    '"with_unknown=True"')
                idx.append(ci)
            out.append(idx)
        return out

    def str2tensor(self, strings):
        padded = []
        for index in self.str2idx(strings):
            src = torch.LongTensor([self.start_idx] + index + [self.end_idx])
            t = torch.full((self.max_seq_len,), self.padding_idx, dtype=torch.long)
            if src.numel() > self.max_seq_len:
                t = src[: self.max_seq_len]
            else:
                t[: src.numel()] = src
            padded.append(t)
        return torch.stack(padded, 0).long()

    def idx2str(self, indexes):
        return [''.join(self.idx2char[i] for i in index) for index in indexes]

    def tensor2idx(self, outputs, img_metas=None):
        indexes, scores = [], []
        for seq in outputs.softmax(dim=-1):
            mv, mi = torch.max(seq, -1)
            si, ss = [], []
            for ci, cs in zip(mi.tolist(), mv.tolist()):
                if ci == self.padding_idx:
                    continue
                if ci == self.end_idx:
                    break
                si.append(ci); ss.append(cs)
            indexes.append(si); scores.append(ss)
        return indexes, scores


class DINO_Finetune(nn.Module):
    """Drop-in for Dino/model/dino_vision.py:135-290.  `config` carries the attributes the reference reads:
    arch, patch_size, drop_path_rate, decoder_{max_seq_len,n_layers,d_embedding,n_head,d_k,d_v,d_model,d_inner}."""

    def __init__(self, config):
        super().__init__()
        self.label_convertor = AttnConvertor(dict_type='DICT90', max_seq_len=config.decoder_max_seq_len, with_unknown=True)
        config.arch = config.arch.replace("deit", "vit")
        archs = {"vit_tiny": vit_tiny, "vit_small": vit_small, "vit_base": vit_base}
        if config.arch not in archs:
            raise NotImplementedError(f"ccd_b200 implements the ViT backbones of CCD, not {config.arch}")
        self.backbone = archs[config.arch](patch_size=config.patch_size, drop_path_rate=config.drop_path_rate)
        embed_dim = self.backbone.embed_dim
        self.encoder = Mlp(in_features=embed_dim, hidden_features=512, out_features=512, act_layer=nn.GELU, drop=0.1)
        config.decoder_num_classes = self.label_convertor.num_classes()
        config.decoder_start_idx = self.label_convertor.start_idx
        config.decoder_padding_idx = self.label_convertor.padding_idx
        self.decoder = NRTRDecoder(n_layers=config.decoder_n_layers, d_embedding=config.decoder_d_embedding,
                                   n_head=config.decoder_n_head, d_k=config.decoder_d_k, d_v=config.decoder_d_v,
                                   d_model=config.decoder_d_model, d_inner=config.decoder_d_inner, n_position=200, dropout=0.1,
                                   num_classes=config.decoder_num_classes, max_seq_len=config.decoder_max_seq_len,
                                   start_idx=config.decoder_start_idx, padding_idx=config.decoder_padding_idx)
        self.loss = TFLoss(ignore_index=self.label_convertor.padding_idx)
        self._enc_b16 = None
        self._enc_cast = ops.ChunkTable()

    def _encoder_bf16(self):
        ps = [self.encoder.fc1.weight, self.encoder.fc2.weight]
        if self._enc_b16 is None or self._enc_b16[0].device != ps[0].device:
            self._enc_b16 = [torch.empty(p.shape, dtype=torch.bfloat16, device=p.device) for p in ps]
        table, n = self._enc_cast.get([p.detach() for p in ps], self._enc_b16, 2)      # every forward (see NRTRDecoder.bf16_weights)
        ops.multi_tensor(ops.MT_CAST_BF16, table, n)
        return self._enc_b16

    def extract_feat(self, img):
        x, out = self.backbone(img)                      # dino_vision.py:199-203
        return x

    def encode(self, img):
        """backbone tokens -> Mlp -> encoder memory bf16 [N*256, 512] (dino_vision.py:214-221)."""
        if not img.is_cuda:
            raise RuntimeError("ccd_b200.DINO_Finetune runs on CUDA (sm_100a) only; there is no CPU path")
        feat = self.extract_feat(img)
        n, s, E = feat.shape
        w1, w2 = self._encoder_bf16()
        p = self.encoder.drop.p if self.training else 0.0
        xb = CastBf16Fn.apply(feat.reshape(n * s, E))
        e = self.encoder
        return FFNFn.apply(xb, w1, w2, e.fc1.weight, e.fc1.bias, e.fc2.weight, e.fc2.bias, None, p, p, self.decoder._seed())

    def forward(self, img, text, return_loss=True, test_speed=False):
        if return_loss:
            return self.forward_train(img, text)
        return self.forward_test(img, test_speed=test_speed)

    def forward_train(self, img, img_metas):
        mem = self.encode(img)
        targets_dict = {'padded_targets': img_metas}
        out_dec, attn = self.decoder(None, mem, targets_dict, train_mode=True)
        return self.loss(out_dec, targets_dict), attn    # the reference returns the last cross-attention map; unused by its caller

    def forward_test(self, img, test_speed=False):
        with torch.no_grad():
            mem = self.encode(img)
            return self.decoder(None, mem, None, train_mode=False, test_speed=test_speed)

    def forward_test_speed(self, img):
        return self.forward_test(img, test_speed=True)
