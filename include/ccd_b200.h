/* ccd_b200 -- C ABI of the B200-native (sm_100a) kernels behind the CCD pretraining hot path.
 *
 * The reference (TongkunGuan/CCD) has no FFI layer: its "plugin boundary" is the Python module surface that
 * train.py:27-35 imports (Dino.model.dino_vision.ABIDINOModel, Dino.loss.Dino_loss.DINOLoss, ...).  The drop-in
 * Python modules of this repository (Dino/, ccd_b200/) call ONLY the entry points declared here, through ctypes
 * (ccd_b200/lib.py).  Each entry point names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers to DEVICE memory + sizes, a cudaStream_t passed as void*; no allocation, no host
 * synchronisation, no global state besides cached TMA descriptors / kernel attributes; re-entrant per stream.
 * Return 0 on success, negative on error (CCD_ERR_*); the Python side turns non-zero into RuntimeError.
 * bf16 = __nv_bfloat16 storage; all matrices row-major.
 */
#ifndef CCD_B200_H
#define CCD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CCD_OK 0
#define CCD_ERR_ARG (-1)
#define CCD_ERR_CUDA (-2)
#define CCD_ERR_TMAP (-3)
#define CCD_ERR_UNSUPPORTED (-4)

/* ---- GEMM epilogues (ccd_gemm_bf16 `epi`) ---- */
#define CCD_EPI_BF16 0  /* out0 bf16 = acc + bias                                                   */
#define CCD_EPI_GELU 1  /* out0 bf16 = acc + bias (may be NULL: inference) ; out1 bf16 = gelu_erf(acc + bias)  (Mlp.fc1+act, vision_transformer.py:59-61) */
#define CCD_EPI_RESID 2 /* out0 f32 = aux_f32 + s[m/256]*(acc + bias); s = DropPath keep-scale (vision_transformer.py:27-36) or NULL;                         (x = x + f(x), vision_transformer.py:109-110) */
#define CCD_EPI_F32 3   /* out0 f32 = acc + bias ; split-K slices accumulate atomically (caller zero-fills)  */
#define CCD_EPI_DGELU 4 /* out0 bf16 = acc * gelu'(aux_bf16)                        (autograd of nn.GELU); out1 (optional, f32 [N], caller zero-fills) += column sums of the fp32 output = bias gradient of the layer below */
#define CCD_EPI_POS 5   /* out0 f32 = acc + bias + aux_f32[(m % 256), :]            (prepare_tokens, vision_transformer.py:225-236) */

/* tcgen05/TMEM/TMA GEMM: C[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulate.
 * a_mn / b_mn = 0: operand stored [rows, K] (K contiguous); = 1: operand stored [K, rows] (read transposed, no copy).
 * Replaces every nn.Linear / Conv2d(k=s=4) contraction and its autograd on the path:
 *   PatchEmbed.proj  vision_transformer.py:126-131 ; Attention.qkv/.proj :82,:90 ; Mlp.fc1/.fc2 :59-65 ;
 *   DINOHead.mlp / last_layer :324-328.   N % 8 == 0; K-extent leading dims % 8 == 0; pointers 16-byte aligned. */
int ccd_gemm_bf16(const void* A, const void* B, int M, int N, int K, int a_mn, int b_mn, int epi, const float* bias,
                  void* out0, void* out1, const void* aux, const float* seq_scale, int ldc, int splits, void* stream);

/* Fused MHSA forward: qkv bf16 [S*256, 3*H*64] (fused Attention.qkv output) -> out bf16 [S*256, H*64],
 * lse2 f32 [S,H,256] (log2-domain row log-sum-exp, for backward; may be NULL).  variant 0: P kept in TMEM,
 * variant 1: P staged through shared memory.  Replaces Attention.forward core, vision_transformer.py:85-89. */
int ccd_mhsa_fwd(const void* qkv, void* out, float* lse2, int S, int H, int variant, void* stream);

/* Fused MHSA backward: -> dqkv bf16 [S*256, 3*H*64]; delta_ws = f32 [2,S,H,256] workspace (rowsum(O*dO) and the negated
 * log-sum-exp in the form the main kernel consumes, written by a small pre-pass).  Replaces autograd of vision_transformer.py:85-89. */
int ccd_mhsa_bwd(const void* qkv, const void* o, const void* d_o, const float* lse2, float* delta_ws, void* dqkv,
                 float* dbias_qkv, const float* dproj_bias, const float* w_proj, int S, int H, void* stream);
/* dbias_qkv (optional, f32 [3*H*64], caller zero-fills): gradient of Attention.qkv.bias = column sums of dqkv.
 *   q part: accumulated by the backward kernel itself from the dQ tiles as they leave TMEM (per-CTA partial sums in shared
 *           memory, one global atomic per (head, column) and CTA);
 *   k part: identically zero (a key bias shifts all scores of a row equally), left untouched;
 *   v part: sum_keys dV = sum_q dO (softmax rows sum to 1) = column sums of d_o.  When the caller passes the gradient of
 *           Attention.proj.bias (dproj_bias f32 [E] = column sums of the proj output gradient) and the proj weight (w_proj f32
 *           [E,E], nn.Linear layout) this is the vector-matrix product dproj_bias . w_proj (d_o = dY . W_proj); with NULLs the
 *           entry point column-sums d_o itself.
 * Pipelined variant only (CCD_ERR_UNSUPPORTED otherwise). */
/* C[M,N] = op(A) B in f32 for the small constant operators of the path: the 256 x 256 bicubic resample of pos_embed
 * (interpolate_pos_encoding, vision_transformer.py:182-201) and its transpose in the backward.  trans_a: A stored [K,M]. */
int ccd_smallmm_f32(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, void* stream);
/* out[c] += sum_r v[r] * W[r, c]   (f32, W row-major [rows, cols]).  On the path: the v part of the Attention.qkv.bias gradient
 * (vision_transformer.py:82) = (Attention.proj.bias gradient) . W_proj, see ccd_mhsa_bwd. */
int ccd_vecmat_add_f32(const float* v, const float* W, float* out, int rows, int cols, void* stream);
/* A/B switch (debug): 1 = pipelined persistent backward kernel (default), 0 = first version (one CTA per (sequence, head)) */
int ccd_set_mhsa_bwd_variant(int value);

/* LayerNorm (eps 1e-6) over f32 rows -> bf16 and/or f32.  Block.norm1/norm2, norm, norm_seg: vision_transformer.py:108-110,247,250 */
int ccd_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, int rows, int E,
                      float eps, void* stream);
/* dx = LN'(dy) + resid; writes f32 and/or bf16 (bf16 copy scaled per sequence by bf16_seq_scale = DropPath of the
 * branch it feeds); dgamma/dbeta accumulate (caller zero-fills); dbias_next (optional) accumulates the column sums of the
 * bf16 copy = bias gradient of the linear layer that consumes it as dY. */
int ccd_layernorm_bwd(const float* x, const float* gamma, const void* dy, int dy_is_bf16, const float* resid, float* dx_f32,
                      void* dx_bf16, float* dgamma, float* dbeta, const float* bf16_seq_scale, float* dbias_next, int rows,
                      int E, float eps, void* stream);

/* out[c] += sum_r x[r,c]  (bias gradients / teacher-centre batch sum Dino/loss/Dino_loss.py:138); out zero-filled by caller */
int ccd_colsum_bf16(const void* x, float* out, int rows, int cols, void* stream);
int ccd_colsum_f32(const float* x, float* out, int rows, int cols, void* stream);

/* F.normalize(dim=-1) (vision_transformer.py:326) and weight_norm of last_layer (:313-316) */
int ccd_l2norm_fwd(const float* x, void* y_bf16, float* inv_norm, int rows, int cols, void* stream);
int ccd_l2norm_bwd(const float* x, const float* inv_norm, const float* dy, void* dx_bf16, int rows, int cols, void* stream);
int ccd_weightnorm_fwd(const float* v, const float* g, void* w_bf16, float* inv_norm, int rows, int cols, void* stream);
int ccd_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int rows,
                       int cols, void* stream);

/* multi-tensor ops over a device chunk table int64[n_chunks][3] = (src_ptr, dst_ptr, n_elems):
 * op 0: dst_bf16 = src_f32 ; op 1: dst = a*dst + b*src (teacher EMA, train.py:264-272) ; op 2: dst = a*src ;
 * op 3: *dst += sum(src^2) ; op 4: dst *= a/(sqrt(*src)+1e-6) if < 1  (per-parameter clip, Dino/modules/utils.py:132-141) */
int ccd_multi_tensor(int op, const void* table_dev, int n_chunks, float a, float b, void* stream);
int ccd_cast_f32_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* Fused optimizer step over a device table int64[n_chunks][10] =
 *   (param, grad_ref, exp_avg, exp_avg_sq, sqnorm_ptr|0, ema_dst|0, param_bf16|0, ema_bf16|0, n_elems, flags);
 *   grad_ref = tensor_index | (element_offset << 16) into grad_ptrs_dev (int64[n_tensors], refreshed per step):
 *   g' = g * min(1, clip/(sqrt(*sqnorm)+1e-6))      per-parameter clip, Dino/modules/utils.py:132-141 (clip <= 0 or ptr 0: off)
 *   AdamW (torch.optim.AdamW as train.py:131-133,252 drives it): p *= 1-lr*wd (flags&1); m,v moments; p -= lr/bc1 * m/(sqrt(v)/sqrt(bc2)+eps)
 *   ema_dst = ema_m*ema_dst + (1-ema_m)*p            teacher EMA, train.py:264-272   (flags&2: EMA only, no AdamW update)
 *   param_bf16 / ema_bf16 = bf16 copies of the updated values (the GEMM operand copies of the next forward).
 * ccd_grad_sqnorm: *sqnorm_ptr += sum(g^2) per row of the same table (caller zero-fills the norms). */
int ccd_grad_sqnorm(const void* table_dev, const void* grad_ptrs_dev, int n_chunks, void* stream);
int ccd_fused_adamw(const void* table_dev, const void* grad_ptrs_dev, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float bias_correction1, float bias_correction2, float clip, float ema_m, void* stream);
/* center = center*m + (sum/denom)*(1-m)   (DINOLoss.update_center, Dino/loss/Dino_loss.py:140-143) */
int ccd_center_ema(float* center, const float* sum, float denom, float momentum, int n, void* stream);
/* x f32 [N,3,32,128] -> bf16 [N*256, 64] patch columns (k = c*16+ky*4+kx, zero padded 48..63) */
int ccd_patch_im2col(const float* x, void* cols_bf16, int n_img, void* stream);

/* DINO cross-view distillation CE (DINOLoss.forward, Dino/loss/Dino_loss.py:81-105).  zs/zt f32 [2R,K];
 * row_loss [2R], stats [2R,4], loss_out[1] = mean over the 2R cross-view pairs. */
int ccd_dino_ce_fwd(const float* zs, const float* zt, const float* center, float student_temp, float teacher_temp,
                    float* row_loss, float* stats, float* loss_out, int R, int K, void* stream);
int ccd_dino_ce_bwd(const float* zs, const float* zt, const float* center, const float* stats, const float* gscale_dev,
                    float student_temp, float teacher_temp, void* dz_bf16, int R, int K, void* stream);
/* Segmentation CE on softmaxed probabilities (Dino_loss.py:63-68,15-26).  logits f32 [N,2,hw], gt f32 [N,hw];
 * partial_ws >= 296 floats. */
int ccd_seg_ce_fwd(const float* logits, const float* gt, float* partial_ws, float* loss_out, int n_img, int hw, void* stream);
int ccd_seg_ce_bwd(const float* logits, const float* gt, const float* gscale_dev, float* dlogits, int n_img, int hw, void* stream);

/* Connected-component character segments (label_cluster.forward, Dino/utils/DBSCAN.py:61-103; loop dino_vision.py:59-71).
 * mode 0: src = f32 masks [n,32,128]; mode 1: src = f32 seg logits [*,2,32,128] (foreground = logit1 > logit0).
 * bits u32 [n,32,128] (bit s = slot s) and/or compact u8 [n,32,128] (slot+1, 0 = background); n_comp int[n]. */
int ccd_ccl_label(const float* src, int mode, void* bits_u32, void* compact_u8, int* n_comp, int n_img, void* stream);
/* Text-mask generation (clusterpixels, mask_create/generate_mask.py:13-29): exact 2-means over the grey levels of each image
 * (histogram + threshold scan) + the border-majority polarity flip.  grey u8 [n,H,W] -> mask f32 {0,1} [n,H,W] (text = 1), the
 * form ccd_ccl_label mode 0 consumes.  H*W <= 2^20. */
int ccd_kmeans_mask(const void* grey_u8, float* mask, int n_img, int H, int W, void* stream);
/* Affine theta of the irregular view from the inverse pixel-space warp matrices (datasetsupervised_kmeans.py:63-71):
 * m_inv f64 [n,3,3], src_hw int32 [n,2] (height, width of each source image) -> theta f32 [n,3,3] for F.affine_grid at img_h x img_w. */
int ccd_affine_theta(const double* m_inv, const int* src_hw, float* theta, int n, int img_h, int img_w, void* stream);
/* affine_grid + bilinear grid_sample(zeros, align_corners=False) + (> 0.1): dino_vision.py:72-77, train.py:234-236 */
int ccd_warp_bits(const void* src_bits, const float* theta, void* dst_bits, int n_img, void* stream);
int ccd_warp_mask(const float* src, const float* theta, float* dst, int n_img, void* stream);
int ccd_bits_to_dense(const void* bits, float* dense, int n_img, void* stream);
int ccd_dense_to_bits(const float* dense, void* bits, int n_img, void* stream);
/* pooling plan: tot4 int[2n,26], cnt int[n], offs int[n+1] (offs[n] = rows per view R), new_index u8 [n,26]
 * (ragged select, dino_vision.py:82-87) */
int ccd_char_plan(const void* bits, int* tot4, int* cnt, int* offs, void* new_index_u8, int n_view, void* stream);
/* mask-guided character pooling (ABIDINOModel.attention, dino_vision.py:38-49) writing only the selected rows [2R,E] */
int ccd_char_pool_fwd(const void* tokens, int tokens_bf16, const void* bits, const int* tot4, const int* cnt, const int* offs,
                      float* rows, int n_view, int E, void* stream);
int ccd_char_pool_bwd(const float* drows, const void* bits, const int* tot4, const int* cnt, const int* offs, float* dtokens,
                      int n_view, int E, void* stream);

/* ---- SegHead (Dino/modules/segmentor.py:37-95): implicit-GEMM convolutions + BatchNorm(train)+ReLU, NHWC bf16 ----
 * ccd_conv_gemm: the persistent tcgen05 GEMM with one operand gathered from an activation {C_total, W, P, H, n_img} through
 * 5-D TMA boxes shifted per tap (zero fill = padding).  spatial_operand 1: C[M=positions, N] = sum_taps shift_tap(sp) * other[N, K]^T
 * (conv3x3 / ConvTranspose2d forward and dgrad; rowmap=1 scatters rows to the (py,px) parity positions of a 2x upsampled output);
 * spatial_operand 2: C[M, N=(tap, channel)] = other[K=positions, M]^T * shift_tap(sp)  (weight gradients, split-K).
 * taps_host = host int[n_taps][4] = (dy, dx, parity plane, channel base); epi = CCD_EPI_BF16 | CCD_EPI_F32. */
int ccd_conv_gemm(const void* sp, const void* other, int M, int N, int K, int epi, const float* bias, void* out0, int ldc,
                  int splits, int spatial_operand, int H, int W, int C_total, int P, int n_img, int n_taps, const int* taps_host,
                  int cols_per_tap, int rowmap, int py, int px, void* stream);
/* BatchNorm2d training statistics: sums[0:C] += sum x, sums[C:2C] += sum x^2 over the M rows of x bf16 [M, C] (ld ldx) */
int ccd_bn_stats(const void* x, int ldx, float* sums_zeroed, int M, int C, void* stream);
/* mean/rstd from the (all-reduced) sums over `count` rows; running_mean/var (nullable) updated with `momentum` */
int ccd_bn_finalize(const float* sums, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                    float* running_var, int C, void* stream);
/* y = relu((x - mean) * rstd * gamma + beta)  -> bf16 [M, C] with leading dimension ldy (torch.cat fused through ldy) */
int ccd_bn_apply_relu(const void* x, int ldx, const float* mean, const float* rstd, const float* gamma, const float* beta, void* y,
                      int ldy, int M, int C, void* stream);
/* BN+ReLU backward pass 1: sums[0:C] += sum dz, sums[C:2C] += sum dz*xhat, dz = dy * [y > 0]; pass 2: dx bf16 */
int ccd_bn_bwd_reduce(const void* dy, int dy_is_f32, int lddy, const void* x, int ldx, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, float* sums_zeroed, int M, int C, void* stream);
int ccd_bn_bwd_apply(const void* dy, int dy_is_f32, int lddy, const void* x, int ldx, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, const float* sums, float inv_m, void* dx, int lddx, int M, int C,
                     void* stream);

/* cls = Conv2d(128 -> 2, 3x3, pad 1) of the SegHead (segmentor.py:88,94) on the [n,32,128,128] bf16 NHWC map: CUDA-core kernels
 * (two output channels cannot fill a tensor-core tile).  w = fp32 [2,128,3,3] (torch layout), logits / dl = fp32 NCHW [n,2,32,128]. */
int ccd_seg_cls_fwd(const void* u2, const float* w, const float* bias, float* logits, int n_img, void* stream);
int ccd_seg_cls_dgrad(const float* dl, const float* w, void* du2_bf16, int n_img, void* stream);
int ccd_seg_cls_wgrad(const void* u2, const float* dl, float* dw_zeroed, float* dbias_zeroed, int n_img, void* stream);

/* ---- recognition / fine-tuning decoder (BASELINE config 5) ----
 * Attention of the NRTR decoder (Dino/decoder/transformer_module.py:9-96): o = dropout(softmax(mask(q k^T / 8))) v per
 * (sample, head), d_k = d_v = 64, tq <= 32 queries, tk <= 256 keys.  q/k/v/o are bf16 matrices [n*tq or n*tk, ld] whose
 * columns [64h, 64h+64) belong to head h (so fused QKV / KV projection outputs are read in place).
 * trg != NULL (self-attention, tk == tq): key j is visible to query i iff j <= i and trg[n, j] != pad_idx
 * (get_pad_mask & get_subsequent_mask, Dino/decoder/nrtr_decoder.py:82-96); trg == NULL: no mask (src_mask is None).
 * p_drop = ScaledDotProductAttention.dropout on the probabilities (0 = evaluation); the mask is a pure function of
 * (seed, sample, head, i, j): the backward regenerates it.  lse [n, heads, tq] (natural log) is kept for the backward. */
int ccd_dec_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                     const long long* trg, int pad_idx, int n, int heads, int tq, int tk, float p_drop, unsigned long long seed,
                     void* stream);
int ccd_dec_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, const void* d_o, int ldo,
                     const float* lse, const long long* trg, int pad_idx, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                     int n, int heads, int tq, int tk, float p_drop, unsigned long long seed, void* stream);

/* A/B switch (process-global): 1 = warp-MMA (mma.sync m16n8k16) decoder attention kernels [default], 0 = scalar CUDA-core kernels */
int ccd_set_dec_attn_variant(int v);

/* TFLoss (Dino/loss/ce_loss.py:94-128): cross-entropy of logits[:, :-1] against targets[:, 1:], pad_idx ignored.
 * logits f32 [n*t, ld] (first n_classes columns valid); acc_zeroed[0] += sum of row losses, acc_zeroed[1] += counted rows
 * (loss = acc[0] / acc[1]); dlogits f32 [n*t, ld] = softmax - onehot on counted rows, 0 elsewhere (scale by g / acc[1]). */
int ccd_tf_ce(const float* logits, int ld, int n_classes, const long long* targets, int n, int t, int pad_idx, float* acc_zeroed,
              float* dlogits, void* stream);

/* nn.Dropout (+ residual): out[i] = resid[i] + keep(seed, i) * x[i] / (1 - p); resid may be NULL; x / out f32 or bf16.
 * The same call with the same seed applied to the output gradient is the backward (Dino/decoder/transformer_module.py:70,
 * :116; Dino/model/dino_vision.py:124). */
int ccd_dropout(const void* x, int x_is_bf16, const float* resid, void* out, int out_is_bf16, long long n, float p,
                unsigned long long seed, void* stream);

/* A/B switch (process-global): 1 = warp-MMA (mma.sync) classifier-convolution kernels [default], 0 = CUDA-core kernels */
int ccd_set_seg_cls_variant(int v);

/* debug / A-B switches (process-global): key 0 = GEMM variant (1 = persistent [default], 0 = one tile per CTA);
   key 1 = epilogue of full tiles in the persistent GEMM (1 = per-shape choice [default], 0 = shared-memory transpose,
   2 = transpose-free thread-per-row wherever alignment allows);
   key 2 = programmatic dependent launch of the GEMM / LayerNorm / attention kernels (0 = off [default: measured 1.6 % slower], 1 = on) */
int ccd_set_option(int key, int value);

/* library identification (build sanity) */
int ccd_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CCD_B200_H */
