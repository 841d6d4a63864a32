"""Thin torch-tensor wrappers over the C ABI (include/ccd_b200.h).  PyTorch is plumbing only here: device memory,
the current CUDA stream and autograd bookkeeping; every computation below is a hand-written sm_100a kernel.
No CPU / eager fallback exists -- a missing library or a non-CUDA tensor raises.
"""
import ctypes
import re

import torch

from . import lib as _lib

EPI_BF16, EPI_GELU, EPI_RESID, EPI_F32, EPI_DGELU, EPI_POS = 0, 1, 2, 3, 4, 5
MT_CAST_BF16, MT_EMA, MT_SCALE, MT_SQNORM, MT_CLIP = 0, 1, 2, 3, 4
LN_EPS = 1e-6
SLOTS = 26

_CTYPE = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
          "unsigned long long": ctypes.c_ulonglong}
_FN = {}
LAUNCHES = [0]     # number of kernels launched through the C ABI (bench.py reports it)
KERNELS_PER_CALL = {"ccd_dino_ce_fwd": 2, "ccd_seg_ce_fwd": 2, "ccd_char_plan": 3, "ccd_mhsa_bwd": 3}
# bench.py roofline instrumentation: when a dict, every call of a listed entry point is bracketed by CUDA events on
# the launching stream:  PROFILE = {"names": {"ccd_gemm_bf16", ...}, "events": []}
PROFILE = None


def _bind():
    if _FN:
        return
    L = _lib.load()
    with open(_lib.HEADER) as f:
        text = f.read()
    for m in re.finditer(r"^int\s+(ccd_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.M | re.S):
        name, args = m.group(1), m.group(2)
        fn = getattr(L, name)
        types = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(ctypes.c_void_p)
                else:
                    types.append(_CTYPE[" ".join(a.split()[:-1]).replace("const ", "").strip()])
        fn.argtypes = types
        fn.restype = ctypes.c_int
        _FN[name] = fn
    import os
    if "CCD_GEMM_VARIANT" in os.environ:            # debug A/B switch (1 = persistent [default], 0 = one tile per CTA)
        _FN["ccd_set_option"](0, int(os.environ["CCD_GEMM_VARIANT"]))
    if "CCD_GEMM_EPILOGUE" in os.environ:           # debug A/B switch (1 = per-shape [default], 0 = smem transpose, 2 = thread-per-row)
        _FN["ccd_set_option"](1, int(os.environ["CCD_GEMM_EPILOGUE"]))
    if "CCD_PDL" in os.environ:                     # debug A/B switch: programmatic dependent launch (0 = off [default], 1 = on)
        _FN["ccd_set_option"](2, int(os.environ["CCD_PDL"]))
    if "CCD_MHSA_BWD_VARIANT" in os.environ:        # 1 = pipelined persistent [default], 0 = first version
        _FN["ccd_set_mhsa_bwd_variant"](int(os.environ["CCD_MHSA_BWD_VARIANT"]))
        _MHSA_BWD_VARIANT[0] = 1 if int(os.environ["CCD_MHSA_BWD_VARIANT"]) else 0


def _call(name, *args, work=None):
    _bind()
    prof = PROFILE
    if prof is not None and name in prof["names"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _FN[name](*args)
        e1.record()
        prof["events"].append((name, work, e0, e1))
    else:
        rc = _FN[name](*args)
    LAUNCHES[0] += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f"{name} failed with code {rc}")


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ccd_b200 ops need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype):
    assert t.dtype == dtype and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


# ------------------------------------------------------------------------------------------------------------
def gemm(A, B, M, N, K, a_mn=0, b_mn=0, epi=EPI_BF16, bias=None, out0=None, out1=None, aux=None, ldc=0, splits=1,
         seq_scale=None):
    """C[M,N] = A[M,K] B[N,K]^T on tcgen05.  See include/ccd_b200.h:ccd_gemm_bf16."""
    _call("ccd_gemm_bf16", _p(A), _p(B), M, N, K, a_mn, b_mn, epi, _p(bias), _p(out0), _p(out1), _p(aux), _p(seq_scale), ldc,
          splits, _s(), work=(2.0 * M * N * K, (M, N, K, a_mn, b_mn, epi)))
    return out0


def gemm_bn(n):
    """N-tile of the persistent GEMM (mirrors dispatch_gemm_persistent in csrc/gemm_umma.cu)."""
    return 256 if n % 256 == 0 else 192 if n % 192 == 0 else 128


def wgrad_splits(n_out, k_in, t_rows):
    """Split-K factor of a weight-gradient GEMM dW[n_out, k_in]: enough (tile, k-slice) work items for two rounds of the
    148 persistent CTAs."""
    bn = gemm_bn(k_in)
    tiles = ((n_out + 127) // 128) * ((k_in + bn - 1) // bn)
    kb = (t_rows + 63) // 64
    return max(1, min(kb, (296 + tiles - 1) // tiles))


def linear_fwd(x_bf16, w_bf16, bias, epi, out0, out1=None, aux=None, seq_scale=None):
    """x [T,K] bf16, w [N,K] bf16 (nn.Linear layout)."""
    T, K = x_bf16.shape
    N = w_bf16.shape[0]
    return gemm(x_bf16, w_bf16, T, N, K, 0, 0, epi, bias, out0, out1, aux, seq_scale=seq_scale)


def linear_dgrad(dy_bf16, w_bf16, epi, out0, aux=None, colsum=None):
    """dx [T,K] = dy [T,N] @ w [N,K]: w is read MN-major (no transpose copy).  EPI_DGELU: `colsum` (optional, zero-filled
    f32 [K]) accumulates the column sums of the fp32 result in the epilogue (bias gradient of the layer below)."""
    T, N = dy_bf16.shape
    K = w_bf16.shape[1]
    return gemm(dy_bf16, w_bf16, T, K, N, 0, 1, epi, None, out0, colsum, aux)


def linear_wgrad(dy_bf16, x_bf16, dw_f32_zeroed):
    """dw [N,K] += dy[T,N]^T @ x[T,K]: both activations read MN-major; split-K over T with fp32 atomics."""
    T, N = dy_bf16.shape
    K = x_bf16.shape[1]
    sp = wgrad_splits(N, K, T)
    return gemm(dy_bf16, x_bf16, N, K, T, 1, 1, EPI_F32, None, dw_f32_zeroed, None, None, dw_f32_zeroed.shape[1], sp)


# 2 = persistent kernel (default); 0 / 1 = one CTA per (sequence, head, query tile) with P in TMEM / in shared memory
MHSA_FWD_VARIANT = int(__import__("os").environ.get("CCD_MHSA_FWD_VARIANT", "2"))


def mhsa_fwd(qkv, S, H, want_lse=True, variant=None):
    if variant is None:
        variant = MHSA_FWD_VARIANT
    E = H * 64
    out = torch.empty(S * 256, E, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(S, H, 256, dtype=torch.float32, device=qkv.device) if want_lse else None
    _call("ccd_mhsa_fwd", _p(_chk(qkv, torch.bfloat16)), _p(out), _p(lse), S, H, variant, _s(),
          work=(4.0 * 256 * 256 * 64 * S * H, (S, H)))
    return out, lse


_MHSA_BWD_VARIANT = [1]


def mhsa_bwd(qkv, o, d_o, lse, S, H, dbias=None, dproj_bias=None, w_proj=None):
    """dqkv of the fused attention; `dbias` (optional, zero-filled f32 [3E]) accumulates the qkv-bias gradient: the q part inside
    the pipelined kernel, the v part (= column sums of d_o) as `dproj_bias @ w_proj` when the caller holds the proj-bias
    gradient (f32 [E]) and the proj weight (f32 [E,E]) -- otherwise by a column-sum pass over d_o; first variant: a column-sum
    pass over dqkv."""
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((2,) + tuple(lse.shape), dtype=torch.float32, device=lse.device)     # workspace [2,S,H,256]
    fused = dbias is not None and _MHSA_BWD_VARIANT[0] == 1
    vm = fused and dproj_bias is not None and w_proj is not None
    _call("ccd_mhsa_bwd", _p(qkv), _p(o), _p(_chk(d_o, torch.bfloat16)), _p(lse), _p(delta), _p(dqkv), _p(dbias) if fused else None,
          _p(_chk(dproj_bias, torch.float32)) if vm else None, _p(_chk(w_proj, torch.float32)) if vm else None,
          S, H, _s(), work=(10.0 * 256 * 256 * 64 * S * H, (S, H)))
    if dbias is not None and not fused:
        colsum_bf16(dqkv, dbias)
    return dqkv


def set_gemm_epilogue(v):
    """Full-tile epilogue of the persistent GEMM: 1 = per-shape choice (default), 0 = shared-memory transpose, 2 = transpose-free
    thread-per-row wherever alignment allows (A/B switch, see include/ccd_b200.h:ccd_set_option)."""
    _bind()
    _FN["ccd_set_option"](1, int(v))


def set_mhsa_bwd_variant(v):
    """1 = pipelined persistent backward kernel (default), 0 = first version (A/B switch, see include/ccd_b200.h)."""
    _bind()
    _FN["ccd_set_mhsa_bwd_variant"](int(v))
    _MHSA_BWD_VARIANT[0] = 1 if int(v) else 0


def layernorm_fwd(x, gamma, beta, want_bf16=True, want_f32=False, eps=LN_EPS):
    rows, E = x.shape
    yb = torch.empty(rows, E, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    yf = torch.empty(rows, E, dtype=torch.float32, device=x.device) if want_f32 else None
    _call("ccd_layernorm_fwd", _p(_chk(x, torch.float32)), _p(gamma), _p(beta), _p(yb), _p(yf), rows, E, eps, _s())
    return yb, yf


def layernorm_bwd(x, gamma, dy, resid, dgamma, dbeta, want_f32=True, want_bf16=True, bf16_seq_scale=None, dbias_next=None,
                  eps=LN_EPS):
    rows, E = x.shape
    assert dy.is_contiguous() and dy.shape == x.shape
    dxf = torch.empty(rows, E, dtype=torch.float32, device=x.device) if want_f32 else None
    dxb = torch.empty(rows, E, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    _call("ccd_layernorm_bwd", _p(x), _p(gamma), _p(dy), 1 if dy.dtype == torch.bfloat16 else 0, _p(resid), _p(dxf), _p(dxb),
          _p(dgamma), _p(dbeta), _p(bf16_seq_scale), _p(dbias_next), rows, E, eps, _s())
    return dxf, dxb


def colsum_bf16(x, out_zeroed):
    _call("ccd_colsum_bf16", _p(_chk(x, torch.bfloat16)), _p(out_zeroed), x.shape[0], x.shape[1], _s())
    return out_zeroed


def smallmm(A, B, trans_a=False):
    """C = op(A) @ B in f32 (small constant operators: the pos-embed resample matrix)."""
    K, N = B.shape
    M = A.shape[1] if trans_a else A.shape[0]
    C = torch.empty(M, N, dtype=torch.float32, device=B.device)
    _call("ccd_smallmm_f32", _p(_chk(A, torch.float32)), _p(_chk(B, torch.float32)), _p(C), M, N, K, 1 if trans_a else 0, _s())
    return C


def vecmat_add(v, W, out):
    """out[c] += sum_r v[r] W[r, c]  (f32)."""
    _call("ccd_vecmat_add_f32", _p(_chk(v, torch.float32)), _p(_chk(W, torch.float32)), _p(out), W.shape[0], W.shape[1], _s())
    return out


def colsum_f32(x, out_zeroed):
    _call("ccd_colsum_f32", _p(_chk(x, torch.float32)), _p(out_zeroed), x.shape[0], x.shape[1], _s())
    return out_zeroed


def l2norm_fwd(x):
    rows, cols = x.shape
    y = torch.empty(rows, cols, dtype=torch.bfloat16, device=x.device)
    inv = torch.empty(rows, dtype=torch.float32, device=x.device)
    _call("ccd_l2norm_fwd", _p(_chk(x, torch.float32)), _p(y), _p(inv), rows, cols, _s())
    return y, inv


def l2norm_bwd(x, inv, dy):
    rows, cols = x.shape
    dx = torch.empty(rows, cols, dtype=torch.bfloat16, device=x.device)
    _call("ccd_l2norm_bwd", _p(x), _p(inv), _p(_chk(dy, torch.float32)), _p(dx), rows, cols, _s())
    return dx


def weightnorm_fwd(v, g, w_bf16=None):
    rows, cols = v.shape
    if w_bf16 is None:
        w_bf16 = torch.empty(rows, cols, dtype=torch.bfloat16, device=v.device)
    inv = torch.empty(rows, dtype=torch.float32, device=v.device)
    _call("ccd_weightnorm_fwd", _p(_chk(v, torch.float32)), _p(g), _p(w_bf16), _p(inv), rows, cols, _s())
    return w_bf16, inv


def weightnorm_bwd(dw, v, g, inv):
    rows, cols = v.shape
    dv = torch.empty_like(v)
    dg = torch.empty(rows, 1, dtype=torch.float32, device=v.device)
    _call("ccd_weightnorm_bwd", _p(_chk(dw, torch.float32)), _p(v), _p(g), _p(inv), _p(dv), _p(dg), rows, cols, _s())
    return dv, dg


def cast_bf16(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _call("ccd_cast_f32_bf16", _p(_chk(src, torch.float32)), _p(dst), src.numel(), _s())
    return dst


def _upload_table(rows, dev):
    """Pointer table -> device without a host stall: pinned staging + asynchronous copy (a pageable torch.tensor(...).to(dev)
    is a synchronous cudaMemcpy that waits for everything queued on the stream)."""
    host = torch.tensor(rows, dtype=torch.int64)
    if torch.device(dev).type != "cuda":          # host-logic unit tests only; every kernel call refuses CPU tensors
        return host, host
    host = host.pin_memory()
    t = host.to(dev, non_blocking=True)
    return t, host          # the pinned buffer must outlive the copy: the caller keeps it


class ChunkTable:
    """Device pointer table for the multi-tensor kernels; built once per distinct set of data_ptrs (cached)."""
    CHUNK = 1 << 16

    def __init__(self):
        self.cache = {}

    def get(self, srcs, dsts, dst_itemsize):
        key = tuple(t.data_ptr() for t in srcs) + tuple(t.data_ptr() for t in dsts)
        hit = self.cache.get(key)
        if hit is None:
            rows = []
            for s, d in zip(srcs, dsts):
                assert s.is_contiguous() and d.is_contiguous() and s.numel() == d.numel()
                n = s.numel()
                for o in range(0, n, self.CHUNK):
                    rows.append((s.data_ptr() + 4 * o, d.data_ptr() + dst_itemsize * o, min(self.CHUNK, n - o)))
            if len(self.cache) >= 8:
                self.cache.clear()
            hit = self.cache[key] = (*_upload_table(rows, srcs[0].device), len(rows))
        return hit[0], hit[2]


class _ClipEntry:
    pass


class ClipTable:
    """Pointer tables for the multi-tensor per-parameter clip: (grad chunk -> &sqnorm[i]) and (&sqnorm[i] -> grad chunk).
    Gradient buffers are re-allocated every step (zero_grad(set_to_none=True)) and the caching allocator alternates
    between a few addresses: the tables are cached per address set instead of being rebuilt (and re-uploaded) per step."""
    CHUNK = 1 << 16

    def __init__(self):
        self.cache = {}

    def get(self, grads):
        key = tuple(g.data_ptr() for g in grads)
        e = self.cache.get(key)
        if e is None:
            dev = grads[0].device
            e = _ClipEntry()
            e.norms = torch.zeros(len(grads), dtype=torch.float32, device=dev)
            base = e.norms.data_ptr()
            a, b = [], []
            for i, g in enumerate(grads):
                assert g.is_contiguous() and g.dtype == torch.float32
                for o in range(0, g.numel(), self.CHUNK):
                    n = min(self.CHUNK, g.numel() - o)
                    a.append((g.data_ptr() + 4 * o, base + 4 * i, n))
                    b.append((base + 4 * i, g.data_ptr() + 4 * o, n))
            e.t_norm, e._h0 = _upload_table(a, dev)
            e.t_clip, e._h1 = _upload_table(b, dev)
            e.n = len(a)
            if len(self.cache) >= 8:
                self.cache.clear()
            self.cache[key] = e
        return e


def clip_per_parameter_(grads, clip, table):
    """In-place per-parameter L2 clip of a list of fp32 gradient tensors in two launches, no host sync."""
    t = table.get(grads)
    t.norms.zero_()
    multi_tensor(MT_SQNORM, t.t_norm, t.n)
    multi_tensor(MT_CLIP, t.t_clip, t.n, float(clip))
    return t.norms


def multi_tensor(op, table, n, a=0.0, b=0.0):
    _call("ccd_multi_tensor", op, _p(table), n, float(a), float(b), _s())


def fused_adamw(table, grad_ptrs, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, clip, ema_m):
    """Per-parameter clip + AdamW + teacher EMA + bf16 operand refresh in one pass.  See include/ccd_b200.h:ccd_fused_adamw."""
    _call("ccd_fused_adamw", _p(table), _p(grad_ptrs), n, float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
          float(bc1), float(bc2), float(clip), float(ema_m), _s())


def grad_sqnorm(table, grad_ptrs, n):
    _call("ccd_grad_sqnorm", _p(table), _p(grad_ptrs), n, _s())


def patch_im2col(x_img):
    n = x_img.shape[0]
    cols = torch.empty(n * 256, 64, dtype=torch.bfloat16, device=x_img.device)
    _call("ccd_patch_im2col", _p(_chk(x_img, torch.float32)), _p(cols), n, _s())
    return cols


# ---- loss ----
def dino_ce_fwd(zs, zt, center, ts, tt):
    R2, K = zs.shape
    row_loss = torch.empty(R2, dtype=torch.float32, device=zs.device)
    stats = torch.empty(R2, 4, dtype=torch.float32, device=zs.device)
    loss = torch.empty(1, dtype=torch.float32, device=zs.device)
    _call("ccd_dino_ce_fwd", _p(_chk(zs, torch.float32)), _p(_chk(zt, torch.float32)), _p(_chk(center, torch.float32)), float(ts),
          float(tt), _p(row_loss), _p(stats), _p(loss), R2 // 2, K, _s())
    return loss, stats


def dino_ce_bwd(zs, zt, center, stats, gscale, ts, tt):
    R2, K = zs.shape
    dz = torch.empty(R2, K, dtype=torch.bfloat16, device=zs.device)
    _call("ccd_dino_ce_bwd", _p(zs), _p(zt), _p(center), _p(stats), _p(gscale), float(ts), float(tt), _p(dz), R2 // 2, K, _s())
    return dz


def seg_ce_fwd(logits, gt):
    n = logits.shape[0]
    hw = logits.shape[2] * logits.shape[3]
    ws = torch.empty(296, dtype=torch.float32, device=logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    _call("ccd_seg_ce_fwd", _p(_chk(logits, torch.float32)), _p(_chk(gt, torch.float32)), _p(ws), _p(loss), n, hw, _s())
    return loss


def seg_ce_bwd(logits, gt, gscale):
    d = torch.empty_like(logits)
    n = logits.shape[0]
    hw = logits.shape[2] * logits.shape[3]
    _call("ccd_seg_ce_bwd", _p(logits), _p(gt), _p(gscale), _p(d), n, hw, _s())
    return d


def center_update(center, teacher_logits, world_size, momentum, all_reduce=None):
    """DINOLoss.update_center (Dino/loss/Dino_loss.py:133-143) incl. the local_rows*world divisor."""
    K = teacher_logits.shape[1]
    s = torch.zeros(K, dtype=torch.float32, device=teacher_logits.device)
    colsum_f32(teacher_logits, s)
    if all_reduce is not None:
        all_reduce(s)
    _call("ccd_center_ema", _p(center), _p(s), float(teacher_logits.shape[0] * world_size), float(momentum), K, _s())


# ---- character segments ----
def ccl_label(src, mode, n_img, want_compact=False):
    dev = src.device
    bits = torch.empty(n_img, 32, 128, dtype=torch.int32, device=dev)
    compact = torch.empty(n_img, 32, 128, dtype=torch.uint8, device=dev) if want_compact else None
    ncomp = torch.empty(n_img, dtype=torch.int32, device=dev)
    _call("ccd_ccl_label", _p(_chk(src, torch.float32)), mode, _p(bits), _p(compact), _p(ncomp), n_img, _s())
    return bits, compact, ncomp


def kmeans_mask(grey_u8):
    """Text masks from grey crops (clusterpixels, mask_create/generate_mask.py:13-29): grey u8 [n,H,W] -> f32 {0,1} [n,H,W]."""
    n, H, W = grey_u8.shape
    out = torch.empty(n, H, W, dtype=torch.float32, device=grey_u8.device)
    _call("ccd_kmeans_mask", _p(_chk(grey_u8, torch.uint8)), _p(out), n, H, W, _s())
    return out


def affine_theta(m_inv, src_hw, img_h=32, img_w=128):
    """Thetas of the irregular view from imgaug's inverse pixel-space matrices (datasetsupervised_kmeans.py:63-71):
    m_inv f64 [n,3,3], src_hw int32 [n,2] (h, w) -> f32 [n,3,3]."""
    n = m_inv.shape[0]
    out = torch.empty(n, 3, 3, dtype=torch.float32, device=m_inv.device)
    _call("ccd_affine_theta", _p(_chk(m_inv, torch.float64)), _p(_chk(src_hw, torch.int32)), _p(out), n, img_h, img_w, _s())
    return out


def warp_bits(bits, theta):
    out = torch.empty_like(bits)
    _call("ccd_warp_bits", _p(bits), _p(_chk(theta, torch.float32)), _p(out), bits.shape[0], _s())
    return out


def warp_mask(mask, theta):
    out = torch.empty_like(mask)
    _call("ccd_warp_mask", _p(_chk(mask, torch.float32)), _p(_chk(theta, torch.float32)), _p(out), mask.shape[0], _s())
    return out


def bits_to_dense(bits):
    n = bits.shape[0]
    dense = torch.empty(n, SLOTS, 32, 128, dtype=torch.float32, device=bits.device)
    _call("ccd_bits_to_dense", _p(bits), _p(dense), n, _s())
    return dense


def dense_to_bits(dense):
    n = dense.shape[0]
    bits = torch.empty(n, 32, 128, dtype=torch.int32, device=dense.device)
    _call("ccd_dense_to_bits", _p(_chk(dense, torch.float32)), _p(bits), n, _s())
    return bits


def char_plan(bits):
    n_view = bits.shape[0] // 2
    dev = bits.device
    tot4 = torch.empty(2 * n_view, SLOTS, dtype=torch.int32, device=dev)
    cnt = torch.empty(n_view, dtype=torch.int32, device=dev)
    offs = torch.empty(n_view + 1, dtype=torch.int32, device=dev)
    new_index = torch.empty(n_view, SLOTS, dtype=torch.uint8, device=dev)
    _call("ccd_char_plan", _p(bits), _p(tot4), _p(cnt), _p(offs), _p(new_index), n_view, _s())
    return tot4, cnt, offs, new_index


def char_pool_fwd(tokens, bits, tot4, cnt, offs, R):
    n_view = bits.shape[0] // 2
    E = tokens.shape[-1]
    rows = torch.empty(2 * R, E, dtype=torch.float32, device=tokens.device)
    _call("ccd_char_pool_fwd", _p(tokens), 1 if tokens.dtype == torch.bfloat16 else 0, _p(bits), _p(tot4), _p(cnt), _p(offs),
          _p(rows), n_view, E, _s())
    return rows


def char_pool_bwd(drows, bits, tot4, cnt, offs, E):
    n_view = bits.shape[0] // 2
    dtok = torch.empty(2 * n_view * 256, E, dtype=torch.float32, device=drows.device)
    _call("ccd_char_pool_bwd", _p(_chk(drows, torch.float32)), _p(bits), _p(tot4), _p(cnt), _p(offs), _p(dtok), n_view, E, _s())
    return dtok


# ---- SegHead: implicit-GEMM convolutions + BatchNorm/ReLU ----
def conv_gemm(sp, other, M, N, K, epi, bias, out0, ldc, splits, spatial_operand, H, W, C_total, P, n_img, taps, cols_per_tap,
              rowmap=0, py=0, px=0):
    """taps = list of (dy, dx, parity_plane, channel_base).  See include/ccd_b200.h:ccd_conv_gemm."""
    arr = (ctypes.c_int * (4 * len(taps)))(*[int(v) for t in taps for v in t])
    _call("ccd_conv_gemm", _p(sp), _p(other), M, N, K, epi, _p(bias), _p(out0), ldc, splits, spatial_operand, H, W, C_total, P,
          n_img, len(taps), ctypes.cast(arr, ctypes.c_void_p), cols_per_tap, rowmap, py, px, _s(), work=(2.0 * M * N * K, ("conv", M, N, K)))
    return out0


def bn_stats(x, ldx, M, C):
    sums = torch.zeros(2 * C, dtype=torch.float32, device=x.device)
    _call("ccd_bn_stats", _p(x), ldx, _p(sums), M, C, _s())
    return sums


def bn_finalize(sums, count, eps, momentum, running_mean, running_var, C):
    mean = torch.empty(C, dtype=torch.float32, device=sums.device)
    rstd = torch.empty(C, dtype=torch.float32, device=sums.device)
    _call("ccd_bn_finalize", _p(sums), float(count), float(eps), float(momentum), _p(mean), _p(rstd), _p(running_mean),
          _p(running_var), C, _s())
    return mean, rstd


def bn_apply_relu(x, ldx, mean, rstd, gamma, beta, y, ldy, M, C):
    _call("ccd_bn_apply_relu", _p(x), ldx, _p(mean), _p(rstd), _p(gamma), _p(beta), _p(y), ldy, M, C, _s())
    return y


def bn_bwd_reduce(dy, lddy, x, ldx, mean, rstd, gamma, beta, M, C):
    sums = torch.zeros(2 * C, dtype=torch.float32, device=x.device)
    _call("ccd_bn_bwd_reduce", _p(dy), 1 if dy.dtype == torch.float32 else 0, lddy, _p(x), ldx, _p(mean), _p(rstd), _p(gamma),
          _p(beta), _p(sums), M, C, _s())
    return sums


def bn_bwd_apply(dy, lddy, x, ldx, mean, rstd, gamma, beta, sums, inv_m, M, C):
    dx = torch.empty(M, C, dtype=torch.bfloat16, device=x.device)
    _call("ccd_bn_bwd_apply", _p(dy), 1 if dy.dtype == torch.float32 else 0, lddy, _p(x), ldx, _p(mean), _p(rstd), _p(gamma),
          _p(beta), _p(sums), float(inv_m), _p(dx), C, M, C, _s())
    return dx


def set_seg_cls_variant(v):
    """1 = tensor-core (mma.sync) classifier-convolution kernels (default), 0 = CUDA-core kernels (A/B switch)."""
    _bind()
    _FN["ccd_set_seg_cls_variant"](int(v))


def seg_cls_fwd(u2, w, bias, n_img):
    logits = torch.empty(n_img, 2, 32, 128, dtype=torch.float32, device=u2.device)
    _call("ccd_seg_cls_fwd", _p(u2), _p(_chk(w, torch.float32)), _p(bias), _p(logits), n_img, _s())
    return logits


def seg_cls_dgrad(dl, w, n_img):
    du2 = torch.empty(n_img * 4096, 128, dtype=torch.bfloat16, device=dl.device)
    _call("ccd_seg_cls_dgrad", _p(_chk(dl, torch.float32)), _p(_chk(w, torch.float32)), _p(du2), n_img, _s())
    return du2


def seg_cls_wgrad(u2, dl, n_img):
    buf = torch.zeros(2 * 128 * 9 + 2, dtype=torch.float32, device=dl.device)
    _call("ccd_seg_cls_wgrad", _p(u2), _p(_chk(dl, torch.float32)), _p(buf), _p(buf[2304:]), n_img, _s())
    return buf[:2304].view(2, 128, 3, 3), buf[2304:]


# ------------------------------------------------------------------------------------------------------------
# recognition / fine-tuning decoder (include/ccd_b200.h: ccd_dec_attn_*, ccd_tf_ce, ccd_dropout)
# ------------------------------------------------------------------------------------------------------------
def dec_attn_fwd(q, k, v, n, heads, tq, tk, trg=None, pad_idx=0, p_drop=0.0, seed=0, want_lse=True):
    """q [n*tq, ldq], k / v [n*tk, ld] bf16 column views (head h = columns [64h, 64h+64)); returns o [n*tq, 64*heads], lse."""
    o = torch.empty(n * tq, heads * 64, dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(n, heads, tq, dtype=torch.float32, device=q.device) if want_lse else None
    _call("ccd_dec_attn_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0), _p(lse), _p(trg),
          pad_idx, n, heads, tq, tk, float(p_drop), int(seed), _s())
    return o, lse


def dec_attn_bwd(q, k, v, o, d_o, lse, dq, dk, dv, n, heads, tq, tk, trg=None, pad_idx=0, p_drop=0.0, seed=0):
    """Writes dq / dk / dv (bf16 column views with the layouts of q / k / v)."""
    assert d_o.stride(0) == o.stride(0)
    _call("ccd_dec_attn_bwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), _p(d_o), o.stride(0), _p(lse),
          _p(trg), pad_idx, _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), n, heads, tq, tk, float(p_drop),
          int(seed), _s())


def set_dec_attn_variant(v):
    """1 = tensor-core (mma.sync) decoder attention kernels (default), 0 = scalar kernels (A/B switch)."""
    _bind()
    _FN["ccd_set_dec_attn_variant"](int(v))


def tf_ce(logits, n_classes, targets, pad_idx):
    """logits f32 [n*t, ld]; returns (acc [2] = loss sum, counted rows; dlogits f32 [n*t, ld] = softmax - onehot)."""
    n, t = targets.shape
    acc = torch.zeros(2, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    _call("ccd_tf_ce", _p(_chk(logits, torch.float32)), logits.shape[1], n_classes, _p(targets), n, t, pad_idx, _p(acc), _p(dl), _s())
    return acc, dl


def dropout(x, p, seed, resid=None, out_dtype=None):
    """out = resid + keep(seed, i) * x / (1 - p)  (nn.Dropout with a regenerable mask; the backward is the same call on dy)."""
    out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    assert x.is_contiguous() and (resid is None or (resid.is_contiguous() and resid.dtype == torch.float32))
    _call("ccd_dropout", _p(x), 1 if x.dtype == torch.bfloat16 else 0, _p(resid), _p(out), 1 if out.dtype == torch.bfloat16 else 0,
          x.numel(), float(p), int(seed), _s())
    return out
