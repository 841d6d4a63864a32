// Recognition (fine-tuning) decoder kernels, SURVEY.md section 8f #1 / BASELINE config 5:
//   * ccd_dec_attn_fwd / _bwd : the NRTR decoder's two attentions (Dino/decoder/transformer_module.py:9-96):
//       masked self-attention over the T <= 32 target tokens (pad + causal mask, nrtr_decoder.py:82-96,103-105) and
//       cross-attention of those T queries over the 256 encoder tokens; softmax(q k^T / 8) v with d_k = d_v = 64,
//       optional dropout on the probabilities (ScaledDotProductAttention.dropout, p = 0.1 in training).
//     Tq <= 32 query rows make this a CUDA-core problem (0.8 MFLOP per (sample, head) against 12 GFLOP of encoder per
//     sample): one CTA per (sample, head), K / V / Q staged once in shared memory (rows padded to 33 words: conflict-free
//     row-per-lane dot products), one warp per query row, probabilities never leave the SM.
//   * ccd_tf_ce : TFLoss (Dino/loss/ce_loss.py:94-128): cross-entropy of outputs[:, :-1] against targets[:, 1:], PAD ignored.
//   * ccd_dropout : nn.Dropout (+ fused residual add) with a counter-based generator: the mask is a pure function of
//     (seed, element index), so the backward pass regenerates it instead of storing it.
#include "ccd_common.cuh"

namespace ccd {

// ---------------------------------------------------------------------------------------------------------
// counter-based uniform generator: 32 mixed bits per (seed, index)  (lowbias32-style avalanche, two rounds keyed by seed)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long idx, uint32_t thresh) {
  const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  const uint32_t h = mix32(mix32(lo ^ (uint32_t)seed) + hi * 0x9E3779B9U + (uint32_t)(seed >> 32));
  return h >= thresh;                       // P(drop) = thresh / 2^32
}
__host__ __device__ inline uint32_t drop_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
}

constexpr int DA_THREADS = 256;
constexpr int DA_WARPS = DA_THREADS / 32;
constexpr int DA_D = 64;                    // d_k = d_v
constexpr int DA_LD = 66;                   // padded row length (bf16 elements) = 33 words
constexpr int DA_MAX_TQ = 32;
constexpr int DA_MAX_TK = 256;

struct DecAttnParams {
  const bf16 *q, *k, *v;                    // [N*Tq, ldq] / [N*Tk, ldk] / [N*Tk, ldv]; head h = columns [64h, 64h+64)
  int ldq, ldk, ldv;
  bf16* o;                                  // [N*Tq, ldo]
  int ldo;
  float* lse;                               // [N, H, Tq] natural-log logsumexp of the masked scaled scores
  const long long* trg;                     // self-attention: [N, Tq] target tokens (pad + causal mask) or NULL (no mask)
  int pad_idx;
  int n, heads, tq, tk;
  float scale;                              // 1 / sqrt(d_k)
  float p_drop;
  unsigned long long seed;
  // backward only
  const bf16* d_o;
  bf16 *dq, *dk, *dv;                       // same layouts as q / k / v
  int lddq, lddk, lddv;
};

__device__ __forceinline__ void stage_rows(bf16* dst, const bf16* src, int rows, int ld, int tid) {
  // rows x 64 bf16 (128 B per row, 16-byte aligned) -> shared rows of DA_LD elements
  for (int i = tid; i < rows * 8; i += DA_THREADS) {
    const int r = i >> 3, c = i & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + c * 8);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * DA_LD + c * 8);
    d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
  }
}
// row . x : `row` = padded bf16 row in shared memory (lanes read different rows: 33-word stride, conflict-free),
// x = 64 floats held in registers by every lane
__device__ __forceinline__ float dot64(const bf16* row, const float (&x)[DA_D]) {
  float a0 = 0.f, a1 = 0.f;
  const uint32_t* r2 = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const uint32_t u = r2[d];
    a0 = fmaf(bf16lo(u), x[2 * d], a0);
    a1 = fmaf(bf16hi(u), x[2 * d + 1], a1);
  }
  return a0 + a1;
}
// every lane gets the whole 64-element row (scaled) of a [.., ld] bf16 matrix into registers (warp-wide broadcast loads)
__device__ __forceinline__ void load_row64(const bf16* row, float scale, float (&x)[DA_D]) {
  const uint4* r4 = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = __ldg(r4 + c);
    x[8 * c + 0] = bf16lo(u.x) * scale; x[8 * c + 1] = bf16hi(u.x) * scale;
    x[8 * c + 2] = bf16lo(u.y) * scale; x[8 * c + 3] = bf16hi(u.y) * scale;
    x[8 * c + 4] = bf16lo(u.z) * scale; x[8 * c + 5] = bf16hi(u.z) * scale;
    x[8 * c + 6] = bf16lo(u.w) * scale; x[8 * c + 7] = bf16hi(u.w) * scale;
  }
}

// dynamic shared memory: K [tk][66] bf16 | V [tk][66] bf16 | per-warp probabilities [tk] f32
__global__ void __launch_bounds__(DA_THREADS, 2) dec_attn_fwd_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  bf16* sK = reinterpret_cast<bf16*>(da_smem);
  bf16* sV = sK + p.tk * DA_LD;
  float* sP = reinterpret_cast<float*>(sV + p.tk * DA_LD);
  const int n = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, p.ldk, threadIdx.x);
  stage_rows(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, p.ldv, threadIdx.x);
  __syncthreads();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int njj = (p.tk + 31) >> 5;
  float* myp = sP + warp * p.tk;
  for (int i = warp; i < p.tq; i += DA_WARPS) {
    float qr[DA_D];
    load_row64(p.q + ((size_t)n * p.tq + i) * p.ldq + h * DA_D, p.scale, qr);     // (q / temperature) as in the reference
    float s[DA_MAX_TK / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < DA_MAX_TK / 32; ++jj) {
      s[jj] = -INFINITY;
      if (jj < njj) {
        const int j = jj * 32 + lane;
        if (j < p.tk) {
          bool vis = true;
          if (p.trg != nullptr) vis = (j <= i) && (p.trg[(size_t)n * p.tq + j] != (long long)p.pad_idx);
          if (vis) s[jj] = dot64(sK + j * DA_LD, qr);
        }
        mx = fmaxf(mx, s[jj]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < DA_MAX_TK / 32; ++jj) {
      s[jj] = (s[jj] == -INFINITY) ? 0.f : __expf(s[jj] - mx);
      sum += s[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < DA_MAX_TK / 32; ++jj) {
      const int j = jj * 32 + lane;
      if (jj < njj && j < p.tk) {
        float pr = s[jj] * inv;
        if (p.p_drop > 0.f) {
          const unsigned long long idx = (((unsigned long long)blockIdx.x * p.tq + i) << 8) + j;
          pr = keep_elem(p.seed, idx, thresh) ? pr * inv_keep : 0.f;
        }
        myp[j] = pr;
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;                 // lane owns output dims 2*lane, 2*lane+1 (two partial chains)
    int j = 0;
    for (; j + 1 < p.tk; j += 2) {
      const float p0 = myp[j], p1 = myp[j + 1];
      const uint32_t u0 = reinterpret_cast<const uint32_t*>(sV + j * DA_LD)[lane];
      const uint32_t u1 = reinterpret_cast<const uint32_t*>(sV + (j + 1) * DA_LD)[lane];
      o0 = fmaf(p0, bf16lo(u0), o0); o1 = fmaf(p0, bf16hi(u0), o1);
      o2 = fmaf(p1, bf16lo(u1), o2); o3 = fmaf(p1, bf16hi(u1), o3);
    }
    if (j < p.tk) {
      const float p0 = myp[j];
      const uint32_t u0 = reinterpret_cast<const uint32_t*>(sV + j * DA_LD)[lane];
      o0 = fmaf(p0, bf16lo(u0), o0); o1 = fmaf(p0, bf16hi(u0), o1);
    }
    reinterpret_cast<uint32_t*>(p.o + ((size_t)n * p.tq + i) * p.ldo + h * DA_D)[lane] = pack_bf16x2(o0 + o2, o1 + o3);
    if (lane == 0 && p.lse != nullptr) p.lse[((size_t)n * p.heads + h) * p.tq + i] = mx + __logf(sum);
    __syncwarp();
  }
}

// backward: dynamic shared memory: K | V | Q [tq][66] | dO [tq][66] (bf16) | P~ [tq][tk] bf16 | dS [tq][tk] bf16
__global__ void __launch_bounds__(DA_THREADS, 2) dec_attn_bwd_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  bf16* sK = reinterpret_cast<bf16*>(da_smem);
  bf16* sV = sK + p.tk * DA_LD;
  bf16* sQ = sV + p.tk * DA_LD;
  bf16* sDO = sQ + p.tq * DA_LD;
  bf16* sP = sDO + p.tq * DA_LD;
  bf16* sDS = sP + p.tq * p.tk;
  const int n = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, p.ldk, threadIdx.x);
  stage_rows(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, p.ldv, threadIdx.x);
  stage_rows(sQ, p.q + (size_t)n * p.tq * p.ldq + h * DA_D, p.tq, p.ldq, threadIdx.x);
  stage_rows(sDO, p.d_o + (size_t)n * p.tq * p.ldo + h * DA_D, p.tq, p.ldo, threadIdx.x);
  __syncthreads();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int njj = (p.tk + 31) >> 5;
  // ---- phase 1: per query row: P, dP, dS; dQ ----
  for (int i = warp; i < p.tq; i += DA_WARPS) {
    const uint32_t ud = reinterpret_cast<const uint32_t*>(sDO + i * DA_LD)[lane];
    const uint32_t uo = reinterpret_cast<const uint32_t*>(p.o + ((size_t)n * p.tq + i) * p.ldo + h * DA_D)[lane];
    float delta = bf16lo(ud) * bf16lo(uo) + bf16hi(ud) * bf16hi(uo);        // delta_i = dO_i . O_i
    delta = warp_sum(delta);
    const float lse = p.lse[((size_t)n * p.heads + h) * p.tq + i];
    float pr[DA_MAX_TK / 32];
    {
      float x[DA_D];
      load_row64(p.q + ((size_t)n * p.tq + i) * p.ldq + h * DA_D, p.scale, x);
#pragma unroll
      for (int jj = 0; jj < DA_MAX_TK / 32; ++jj) {
        pr[jj] = 0.f;
        const int j = jj * 32 + lane;
        if (jj < njj && j < p.tk) {
          bool vis = true;
          if (p.trg != nullptr) vis = (j <= i) && (p.trg[(size_t)n * p.tq + j] != (long long)p.pad_idx);
          if (vis) pr[jj] = __expf(dot64(sK + j * DA_LD, x) - lse);
        }
      }
    }
    {
      float x[DA_D];
      load_row64(p.d_o + ((size_t)n * p.tq + i) * p.ldo + h * DA_D, 1.0f, x);
#pragma unroll
      for (int jj = 0; jj < DA_MAX_TK / 32; ++jj) {
        const int j = jj * 32 + lane;
        if (jj < njj && j < p.tk) {
          float pt = pr[jj], ds = 0.f;
          if (pt != 0.f) {
            float dp = dot64(sV + j * DA_LD, x);                             // d(P~)_ij = dO_i . V_j
            if (p.p_drop > 0.f) {
              const unsigned long long idx = (((unsigned long long)blockIdx.x * p.tq + i) << 8) + j;
              const bool kp = keep_elem(p.seed, idx, thresh);
              dp = kp ? dp * inv_keep : 0.f;
              pt = kp ? pt * inv_keep : 0.f;
            }
            ds = pr[jj] * (dp - delta);
          }
          sP[i * p.tk + j] = __float2bfloat16(pt);
          sDS[i * p.tk + j] = __float2bfloat16(ds);
        }
      }
    }
    __syncwarp();
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;                 // dQ_i = scale * sum_j dS_ij K_j (bf16-rounded dS, like the dK path)
    int j = 0;
    for (; j + 1 < p.tk; j += 2) {
      const float d0 = __bfloat162float(sDS[i * p.tk + j]), d1 = __bfloat162float(sDS[i * p.tk + j + 1]);
      const uint32_t u0 = reinterpret_cast<const uint32_t*>(sK + j * DA_LD)[lane];
      const uint32_t u1 = reinterpret_cast<const uint32_t*>(sK + (j + 1) * DA_LD)[lane];
      g0 = fmaf(d0, bf16lo(u0), g0); g1 = fmaf(d0, bf16hi(u0), g1);
      g2 = fmaf(d1, bf16lo(u1), g2); g3 = fmaf(d1, bf16hi(u1), g3);
    }
    if (j < p.tk) {
      const float d0 = __bfloat162float(sDS[i * p.tk + j]);
      const uint32_t u0 = reinterpret_cast<const uint32_t*>(sK + j * DA_LD)[lane];
      g0 = fmaf(d0, bf16lo(u0), g0); g1 = fmaf(d0, bf16hi(u0), g1);
    }
    reinterpret_cast<uint32_t*>(p.dq + ((size_t)n * p.tq + i) * p.lddq + h * DA_D)[lane] = pack_bf16x2((g0 + g2) * p.scale, (g1 + g3) * p.scale);
  }
  __syncthreads();
  // ---- phase 2: dV_j = sum_i P~_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i   (warp = key j, lane = dim pair) ----
  for (int j = warp; j < p.tk; j += DA_WARPS) {
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int i = 0; i < p.tq; ++i) {
      const float pt = __bfloat162float(sP[i * p.tk + j]), ds = __bfloat162float(sDS[i * p.tk + j]);
      const uint32_t ud = reinterpret_cast<const uint32_t*>(sDO + i * DA_LD)[lane];
      const uint32_t uq = reinterpret_cast<const uint32_t*>(sQ + i * DA_LD)[lane];
      v0 = fmaf(pt, bf16lo(ud), v0); v1 = fmaf(pt, bf16hi(ud), v1);
      k0 = fmaf(ds, bf16lo(uq), k0); k1 = fmaf(ds, bf16hi(uq), k1);
    }
    reinterpret_cast<uint32_t*>(p.dv + ((size_t)n * p.tk + j) * p.lddv + h * DA_D)[lane] = pack_bf16x2(v0, v1);
    reinterpret_cast<uint32_t*>(p.dk + ((size_t)n * p.tk + j) * p.lddk + h * DA_D)[lane] = pack_bf16x2(k0 * p.scale, k1 * p.scale);
  }
}

// ---------------------------------------------------------------------------------------------------------
// TFLoss: one warp per (sample, position) row of logits [N*T, ld] (first C columns valid)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tf_ce_kernel(const float* __restrict__ logits, int ld, int C, const long long* __restrict__ targets,
                                                    int N, int T, int pad_idx, float* __restrict__ acc /* [2]: loss sum, count */,
                                                    float* __restrict__ dlogits /* [N*T, ld] = softmax - onehot on counted rows, else 0 */) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N * T) return;
  const int n = row / T, t = row - n * T;
  long long tgt = pad_idx;
  if (t + 1 < T) tgt = targets[(size_t)n * T + t + 1];            // outputs[:, :-1] against targets[:, 1:]
  const bool counted = (tgt != (long long)pad_idx);
  const float* z = logits + (size_t)row * ld;
  float* dz = dlogits + (size_t)row * ld;
  if (!counted) {
    for (int c = lane; c < ld; c += 32) dz[c] = 0.f;
    return;
  }
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += __expf(z[c] - mx);
  sum = warp_sum(sum);
  const float lse = mx + __logf(sum);
  for (int c = lane; c < ld; c += 32) dz[c] = (c < C) ? __expf(z[c] - lse) - ((long long)c == tgt ? 1.f : 0.f) : 0.f;
  if (lane == 0) {
    atomicAdd(acc, lse - z[tgt]);
    atomicAdd(acc + 1, 1.0f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// dropout (+ residual): out = resid + keep(seed, i) * x / (1 - p)
// ---------------------------------------------------------------------------------------------------------
template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(256) dropout_kernel(const TIN* __restrict__ x, const float* __restrict__ resid, TOUT* __restrict__ out,
                                                      unsigned long long n, float p, unsigned long long seed) {
  const uint32_t thresh = drop_threshold(p);
  const float inv_keep = 1.0f / (1.0f - p);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    float v;
    if constexpr (sizeof(TIN) == 2) v = __bfloat162float(x[i]); else v = x[i];
    v = keep_elem(seed, i, thresh) ? v * inv_keep : 0.f;
    if (resid != nullptr) v += resid[i];
    if constexpr (sizeof(TOUT) == 2) out[i] = __float2bfloat16(v); else out[i] = v;
  }
}

static size_t dec_attn_smem(int tq, int tk, bool bwd) {
  size_t s = (size_t)2 * tk * DA_LD * 2;
  if (!bwd) return s + (size_t)DA_WARPS * tk * 4;
  return s + (size_t)2 * tq * DA_LD * 2 + (size_t)2 * tq * tk * 2;
}

}  // namespace ccd

using namespace ccd;

// C ABI -- see include/ccd_b200.h
extern "C" int ccd_dec_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                                const long long* trg, int pad_idx, int n, int heads, int tq, int tk, float p_drop,
                                unsigned long long seed, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!q || !k || !v || !o || n <= 0 || heads <= 0 || tq <= 0 || tq > DA_MAX_TQ || tk <= 0 || tk > DA_MAX_TK || p_drop < 0.f || p_drop >= 1.f)
    return CCD_ERR_ARG;
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 1) || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15)) return CCD_ERR_ARG;
  if (trg != nullptr && tk != tq) return CCD_ERR_ARG;             // the target mask is defined for self-attention only
  DecAttnParams p{};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.o = (bf16*)o; p.ldo = ldo; p.lse = lse; p.trg = trg; p.pad_idx = pad_idx;
  p.n = n; p.heads = heads; p.tq = tq; p.tk = tk; p.scale = 0.125f; p.p_drop = p_drop; p.seed = seed;
  const size_t smem = dec_attn_smem(tq, tk, false);
  static bool attr = false;
  if (!attr) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_smem(DA_MAX_TQ, DA_MAX_TK, false)));
    attr = true;
  }
  dec_attn_fwd_kernel<<<n * heads, DA_THREADS, smem, stream>>>(p);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_dec_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, const void* d_o,
                                int ldo, const float* lse, const long long* trg, int pad_idx, void* dq, int lddq, void* dk, int lddk,
                                void* dv, int lddv, int n, int heads, int tq, int tk, float p_drop, unsigned long long seed,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv || n <= 0 || heads <= 0 || tq <= 0 || tq > DA_MAX_TQ || tk <= 0 ||
      tk > DA_MAX_TK || p_drop < 0.f || p_drop >= 1.f)
    return CCD_ERR_ARG;
  if ((ldq & 7) || (ldk & 7) || (ldv & 7) || (ldo & 7) || (lddq & 1) || (lddk & 1) || (lddv & 1) || ((uintptr_t)q & 15) ||
      ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)d_o & 15))
    return CCD_ERR_ARG;
  if (trg != nullptr && tk != tq) return CCD_ERR_ARG;
  DecAttnParams p{};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.o = (bf16*)const_cast<void*>(o); p.ldo = ldo; p.lse = const_cast<float*>(lse); p.trg = trg; p.pad_idx = pad_idx;
  p.n = n; p.heads = heads; p.tq = tq; p.tk = tk; p.scale = 0.125f; p.p_drop = p_drop; p.seed = seed;
  p.d_o = (const bf16*)d_o; p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  const size_t smem = dec_attn_smem(tq, tk, true);
  static bool attr = false;
  if (!attr) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_smem(DA_MAX_TQ, DA_MAX_TK, true)));
    attr = true;
  }
  dec_attn_bwd_kernel<<<n * heads, DA_THREADS, smem, stream>>>(p);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_tf_ce(const float* logits, int ld, int n_classes, const long long* targets, int n, int t, int pad_idx,
                         float* acc_zeroed, float* dlogits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!logits || !targets || !acc_zeroed || !dlogits || n <= 0 || t <= 1 || n_classes <= 0 || ld < n_classes) return CCD_ERR_ARG;
  tf_ce_kernel<<<(n * t + 7) / 8, 256, 0, stream>>>(logits, ld, n_classes, targets, n, t, pad_idx, acc_zeroed, dlogits);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_dropout(const void* x, int x_is_bf16, const float* resid, void* out, int out_is_bf16, long long n, float p,
                           unsigned long long seed, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !out || n <= 0 || p < 0.f || p >= 1.f) return CCD_ERR_ARG;
  const int grid = (int)((n + 256 * 8 - 1) / (256 * 8) < 148 * 16 ? (n + 256 * 8 - 1) / (256 * 8) : 148 * 16);
  const unsigned long long un = (unsigned long long)n;
  if (x_is_bf16 && out_is_bf16) dropout_kernel<bf16, bf16><<<grid, 256, 0, stream>>>((const bf16*)x, resid, (bf16*)out, un, p, seed);
  else if (x_is_bf16) dropout_kernel<bf16, float><<<grid, 256, 0, stream>>>((const bf16*)x, resid, (float*)out, un, p, seed);
  else if (out_is_bf16) dropout_kernel<float, bf16><<<grid, 256, 0, stream>>>((const float*)x, resid, (bf16*)out, un, p, seed);
  else dropout_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)x, resid, (float*)out, un, p, seed);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
