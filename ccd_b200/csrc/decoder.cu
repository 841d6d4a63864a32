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
// Tensor-core variant (default): the same two kernels on warp-level mma.sync m16n8k16 bf16 tiles.  The query dimension
// (T <= 32) is two 16-row tiles, far below a tcgen05 128-row tile, so the legacy warp MMA is the right instrument here:
// per (sample, head) the contractions are 5 MFLOP backward, and the scalar kernels above were shared-memory-bandwidth
// bound at one LDS per two FMAs (11 TFLOP/s).  Shared rows are padded by 8 elements (row stride = 16 B mod 128 B) so that
// every ldmatrix phase touches 8 distinct 16-byte bank groups.
// ---------------------------------------------------------------------------------------------------------
constexpr int MM_LDK = 72;                  // K / V / Q / dO rows: 64 + 8 bf16
constexpr int MM_LDP = 264;                 // P~ / dS rows: 256 + 8 bf16

// A tile 16(m) x 16(k) of a row-major matrix S[m][k]
__device__ __forceinline__ void ld_a(uint32_t (&r)[4], const bf16* S, int ld, int m0, int k0, int lane) {
  ldsm_x4(r, S + (m0 + (lane & 15)) * ld + k0 + (lane >> 4) * 8);
}
// A tile 16(m) x 16(k) of A = X^T where X[k][m] is row-major
__device__ __forceinline__ void ld_a_t(uint32_t (&r)[4], const bf16* X, int ld, int k0, int m0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4_t(r, X + (k0 + (lane & 7) + (mi >> 1) * 8) * ld + m0 + (mi & 1) * 8);
}
// B for two n-tiles (n0 .. n0+15) x 16(k) from Y[n][k] row-major: r[0],r[1] = n-tile 0; r[2],r[3] = n-tile 1
__device__ __forceinline__ void ld_b_nk(uint32_t (&r)[4], const bf16* Y, int ld, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4(r, Y + (n0 + (lane & 7) + (mi >> 1) * 8) * ld + k0 + (mi & 1) * 8);
}
// B for two n-tiles (n0 .. n0+15) x 16(k) from Z[k][n] row-major
__device__ __forceinline__ void ld_b_kn(uint32_t (&r)[4], const bf16* Z, int ld, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4_t(r, Z + (k0 + (lane & 7) + (mi & 1) * 8) * ld + n0 + (mi >> 1) * 8);
}
// rows x 64 bf16 from global (ld elements apart) -> shared rows of MM_LDK elements; rows [rows, rows_pad) are zero-filled.
// Asynchronous 16-byte copies (LDGSTS): a thread has all of its chunks of K, V, Q (and dO) in flight at once instead of one
// load -> store round trip per chunk; stage_commit_wait() precedes the __syncthreads that publishes the tiles.
__device__ __forceinline__ void stage_rows_mm(bf16* dst, const bf16* src, int rows, int rows_pad, int ld, int tid) {
  for (int i = tid; i < rows_pad * 8; i += DA_THREADS) {
    const int r = i >> 3, c = i & 7;
    bf16* d = dst + r * MM_LDK + c * 8;
    if (r < rows) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(src + (size_t)r * ld + c * 8) : "memory");
    } else {
      *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
__device__ __forceinline__ void stage_commit_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ bool key_visible(const DecAttnParams& p, int n, int row, int key) {
  if (row >= p.tq || key >= p.tk) return false;
  if (p.trg == nullptr) return true;
  return key <= row && p.trg[(size_t)n * p.tq + key] != (long long)p.pad_idx;
}

// S fragments of this warp's 32 keys (keys kb .. kb+31) for all 32 query rows: acc[mt][nt][4]
__device__ __forceinline__ void qk_block(float (&acc)[2][4][4], const bf16* sA, const bf16* sB, int kb, int lane) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0[4], a1[4];
    ld_a(a0, sA, MM_LDK, 0, ks * 16, lane);
    ld_a(a1, sA, MM_LDK, 16, ks * 16, lane);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ld_b_nk(b, sB, MM_LDK, kb + np * 16, ks * 16, lane);
      mma_bf16(acc[0][2 * np], a0, b[0], b[1]);
      mma_bf16(acc[0][2 * np + 1], a0, b[2], b[3]);
      mma_bf16(acc[1][2 * np], a1, b[0], b[1]);
      mma_bf16(acc[1][2 * np + 1], a1, b[2], b[3]);
    }
  }
}

// forward.  dynamic shared memory: K [tkp][72] | V [tkp][72] | Q [32][72] | P~ [32][264] (bf16) | red [8][32] f32 | stat [32] f32
__global__ void __launch_bounds__(DA_THREADS, 2) dec_attn_fwd_mma_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int tkp = (p.tk + 31) & ~31;
  bf16* sK = reinterpret_cast<bf16*>(da_smem);
  bf16* sV = sK + tkp * MM_LDK;
  bf16* sQ = sV + tkp * MM_LDK;
  bf16* sP = sQ + 32 * MM_LDK;
  float* sRed = reinterpret_cast<float*>(sP + 32 * MM_LDP);
  float* sStat = sRed + DA_WARPS * 32;
  const int n = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows_mm(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, tkp, p.ldk, threadIdx.x);
  stage_rows_mm(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, tkp, p.ldv, threadIdx.x);
  stage_rows_mm(sQ, p.q + (size_t)n * p.tq * p.ldq + h * DA_D, p.tq, 32, p.ldq, threadIdx.x);
  stage_commit_wait();
  __syncthreads();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int kb = warp * 32;
  const bool active = kb < p.tk;
  const int r0 = lane >> 2, c0 = 2 * (lane & 3);
  float acc[2][4][4];
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // rows r0, r0+8, r0+16, r0+24
  if (active) {
    qk_block(acc, sQ, sK, kb, lane);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = mt * 16 + r0 + (e >> 1) * 8, key = kb + nt * 8 + c0 + (e & 1);
          const float s = key_visible(p, n, row, key) ? acc[mt][nt][e] * p.scale : -INFINITY;
          acc[mt][nt][e] = s;
          mx[mt * 2 + (e >> 1)] = fmaxf(mx[mt * 2 + (e >> 1)], s);
        }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    if ((lane & 3) == 0) sRed[warp * 32 + r0 + 8 * i] = mx[i];
  }
  __syncthreads();
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) m = fmaxf(m, sRed[w * 32 + r0 + 8 * i]);
    mx[i] = m;
  }
  __syncthreads();                                                // sRed is reused for the sums
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = mt * 2 + (e >> 1);
          const float s = acc[mt][nt][e];
          const float pr = (s == -INFINITY) ? 0.f : __expf(s - mx[i]);
          acc[mt][nt][e] = pr;
          sum[i] += pr;
        }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
    if ((lane & 3) == 0) sRed[warp * 32 + r0 + 8 * i] = sum[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) t += sRed[w * 32 + r0 + 8 * i];
    sum[i] = t;
  }
  if (warp == 0 && (lane & 3) == 0 && p.lse != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = r0 + 8 * i;
      if (row < p.tq) p.lse[((size_t)n * p.heads + h) * p.tq + row] = mx[i] + __logf(sum[i]);
    }
  }
  if (active) {                                                   // P~ (dropout applied) -> shared, bf16
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int i = mt * 2 + hf, row = mt * 16 + r0 + hf * 8, key = kb + nt * 8 + c0;
          const float inv = sum[i] > 0.f ? 1.0f / sum[i] : 0.f;
          float p0 = acc[mt][nt][2 * hf] * inv, p1 = acc[mt][nt][2 * hf + 1] * inv;
          if (p.p_drop > 0.f) {
            const unsigned long long idx = (((unsigned long long)blockIdx.x * p.tq + row) << 8) + key;
            p0 = keep_elem(p.seed, idx, thresh) ? p0 * inv_keep : 0.f;
            p1 = keep_elem(p.seed, idx + 1, thresh) ? p1 * inv_keep : 0.f;
          }
          *reinterpret_cast<uint32_t*>(sP + row * MM_LDP + key) = pack_bf16x2(p0, p1);
        }
  }
  __syncthreads();
  // O[32 x 64] = P~[32 x tkp] V[tkp x 64]: warp -> (row tile, 16 output columns)
  {
    const int mt = warp & 1, n0 = (warp >> 1) * 16;
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int kk = 0; kk < tkp; kk += 16) {
      uint32_t a[4], b[4];
      ld_a(a, sP, MM_LDP, mt * 16, kk, lane);
      ld_b_kn(b, sV, MM_LDK, kk, n0, lane);
      mma_bf16(o[0], a, b[0], b[1]);
      mma_bf16(o[1], a, b[2], b[3]);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int row = mt * 16 + r0 + hf * 8;
        if (row < p.tq)
          *reinterpret_cast<uint32_t*>(p.o + ((size_t)n * p.tq + row) * p.ldo + h * DA_D + n0 + t * 8 + c0) = pack_bf16x2(o[t][2 * hf], o[t][2 * hf + 1]);
      }
  }
}

// backward.  dynamic shared memory: K [tkp][72] | V [tkp][72] (later P~ | dS [32][264] each) | Q [32][72] | dO [32][72] | lse, delta [32] f32
__global__ void __launch_bounds__(DA_THREADS, 2) dec_attn_bwd_mma_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int tkp = (p.tk + 31) & ~31;
  bf16* sK = reinterpret_cast<bf16*>(da_smem);
  bf16* sV = sK + tkp * MM_LDK;
  const int v_elems = max(tkp * MM_LDK, 2 * 32 * MM_LDP);        // the V region is reused for P~ and dS
  bf16* sQ = sV + v_elems;
  bf16* sDO = sQ + 32 * MM_LDK;
  float* sLse = reinterpret_cast<float*>(sDO + 32 * MM_LDK);
  float* sDelta = sLse + 32;
  bf16* sP = sV;
  bf16* sDS = sV + 32 * MM_LDP;
  const int n = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows_mm(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, tkp, p.ldk, threadIdx.x);
  stage_rows_mm(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, tkp, p.ldv, threadIdx.x);
  stage_rows_mm(sQ, p.q + (size_t)n * p.tq * p.ldq + h * DA_D, p.tq, 32, p.ldq, threadIdx.x);
  stage_rows_mm(sDO, p.d_o + (size_t)n * p.tq * p.ldo + h * DA_D, p.tq, 32, p.ldo, threadIdx.x);
  for (int row = warp; row < 32; row += DA_WARPS) {               // delta_i = dO_i . O_i ; lse_i
    float d = 0.f;
    if (row < p.tq) {
      const size_t off = ((size_t)n * p.tq + row) * p.ldo + h * DA_D;
      const uint32_t ud = reinterpret_cast<const uint32_t*>(p.d_o + off)[lane];
      const uint32_t uo = reinterpret_cast<const uint32_t*>(p.o + off)[lane];
      d = bf16lo(ud) * bf16lo(uo) + bf16hi(ud) * bf16hi(uo);
    }
    d = warp_sum(d);
    if (lane == 0) {
      sDelta[row] = d;
      sLse[row] = row < p.tq ? p.lse[((size_t)n * p.heads + h) * p.tq + row] : 0.f;
    }
  }
  stage_commit_wait();
  __syncthreads();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int kb = warp * 32;
  const bool active = kb < p.tk;
  const int r0 = lane >> 2, c0 = 2 * (lane & 3);
  uint32_t pk[2][4][2], dk_[2][4][2];                             // packed bf16 pairs of P~ and dS for this warp's 32 keys
  if (active) {
    float acc[2][4][4];
    float pr[2][4][4];
    qk_block(acc, sQ, sK, kb, lane);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = mt * 16 + r0 + (e >> 1) * 8, key = kb + nt * 8 + c0 + (e & 1);
          pr[mt][nt][e] = key_visible(p, n, row, key) ? __expf(acc[mt][nt][e] * p.scale - sLse[row]) : 0.f;
        }
    qk_block(acc, sDO, sV, kb, lane);                             // d(P~) = dO V^T
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int row = mt * 16 + r0 + hf * 8, key = kb + nt * 8 + c0;
          float pt[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pe = pr[mt][nt][2 * hf + e];
            float dp = acc[mt][nt][2 * hf + e];
            pt[e] = pe;
            if (p.p_drop > 0.f) {
              const unsigned long long idx = (((unsigned long long)blockIdx.x * p.tq + row) << 8) + key + e;
              const bool kp = keep_elem(p.seed, idx, thresh);
              pt[e] = kp ? pe * inv_keep : 0.f;
              dp = kp ? dp * inv_keep : 0.f;
            }
            ds[e] = pe * (dp - sDelta[row]);
          }
          pk[mt][nt][hf] = pack_bf16x2(pt[0], pt[1]);
          dk_[mt][nt][hf] = pack_bf16x2(ds[0], ds[1]);
        }
  }
  __syncthreads();                                                // every warp is done reading V: its region becomes P~ | dS
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int row = mt * 16 + r0 + hf * 8, key = kb + nt * 8 + c0;
          *reinterpret_cast<uint32_t*>(sP + row * MM_LDP + key) = pk[mt][nt][hf];
          *reinterpret_cast<uint32_t*>(sDS + row * MM_LDP + key) = dk_[mt][nt][hf];
        }
  }
  __syncthreads();
  // dQ[32 x 64] = scale * dS[32 x tkp] K[tkp x 64]: warp -> (row tile, 16 output columns)
  {
    const int mt = warp & 1, n0 = (warp >> 1) * 16;
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int kk = 0; kk < tkp; kk += 16) {
      uint32_t a[4], b[4];
      ld_a(a, sDS, MM_LDP, mt * 16, kk, lane);
      ld_b_kn(b, sK, MM_LDK, kk, n0, lane);
      mma_bf16(o[0], a, b[0], b[1]);
      mma_bf16(o[1], a, b[2], b[3]);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int row = mt * 16 + r0 + hf * 8;
        if (row < p.tq)
          *reinterpret_cast<uint32_t*>(p.dq + ((size_t)n * p.tq + row) * p.lddq + h * DA_D + n0 + t * 8 + c0) =
              pack_bf16x2(o[t][2 * hf] * p.scale, o[t][2 * hf + 1] * p.scale);
      }
  }
  // dV[keys x 64] = P~^T dO ; dK[keys x 64] = scale * dS^T Q: warp -> its 32 keys (two 16-key tiles), all 64 columns
  if (active) {
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const bf16* sA = which == 0 ? sP : sDS;
      const bf16* sB = which == 0 ? sDO : sQ;
      bf16* dst = which == 0 ? p.dv : p.dk;
      const int ldd = which == 0 ? p.lddv : p.lddk;
      const float sc = which == 0 ? 1.0f : p.scale;
#pragma unroll 1
      for (int kt = 0; kt < 2; ++kt) {
        const int key0 = kb + kt * 16;
        if (key0 >= p.tk) break;
        float o[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
          for (int e = 0; e < 4; ++e) o[t][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t a[4];
          ld_a_t(a, sA, MM_LDP, ks * 16, key0, lane);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ld_b_kn(b, sB, MM_LDK, ks * 16, np * 16, lane);
            mma_bf16(o[2 * np], a, b[0], b[1]);
            mma_bf16(o[2 * np + 1], a, b[2], b[3]);
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int key = key0 + r0 + hf * 8;
            if (key < p.tk)
              *reinterpret_cast<uint32_t*>(dst + ((size_t)n * p.tk + key) * ldd + h * DA_D + t * 8 + c0) =
                  pack_bf16x2(o[t][2 * hf] * sc, o[t][2 * hf + 1] * sc);
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Self-attention over <= 32 target tokens: one WARP per head, one CTA per sample (8 heads).  With one CTA per (sample, head)
// the 4096 tiny CTAs of a batch-512 launch were launch-bound (87 us forward / 151 us backward for 0.2 GFLOP); here a warp owns
// its head end to end -- staging (cp.async), S = QK^T, softmax (quad shuffles only), P~V, and in the backward dP, dS, dQ, dV, dK --
// with no block-level barrier at all.
// ---------------------------------------------------------------------------------------------------------
constexpr int MM_LDS = 40;                  // P~ / dS rows of the warp-per-head kernels: 32 + 8 bf16
constexpr int SW_FWD_BYTES = (3 * 32 * MM_LDK + 32 * MM_LDS) * 2;                    // K, V, Q, P~
constexpr int SW_BWD_BYTES = (4 * 32 * MM_LDK + 2 * 32 * MM_LDS) * 2 + 64 * 4;       // K, V, Q, dO, P~, dS, lse, delta

__device__ __forceinline__ void stage_rows_warp(bf16* dst, const bf16* src, int rows, int ld, int lane) {
  for (int i = lane; i < 32 * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    bf16* d = dst + r * MM_LDK + c * 8;
    if (r < rows) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(src + (size_t)r * ld + c * 8) : "memory");
    } else {
      *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// C[16*MT... ] helper: out[32 x 64] (two row tiles) = A[32 x 32] B[32 x 64] with A row-major (ld_a) or transposed (ld_a_t), B = Z[k][n]
template <bool A_T>
__device__ __forceinline__ void small_mm(float (&o)[8][4], const bf16* sA, int m0, const bf16* sB, int lane) {
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[t][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t a[4];
    if (A_T) ld_a_t(a, sA, MM_LDS, ks * 16, m0, lane);
    else     ld_a(a, sA, MM_LDS, m0, ks * 16, lane);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ld_b_kn(b, sB, MM_LDK, ks * 16, np * 16, lane);
      mma_bf16(o[2 * np], a, b[0], b[1]);
      mma_bf16(o[2 * np + 1], a, b[2], b[3]);
    }
  }
}
__device__ __forceinline__ void store_tile_rows(bf16* dst, int ld, int row0, int n_rows, const float (&o)[8][4], float sc, int lane) {
  const int r0 = lane >> 2, c0 = 2 * (lane & 3);
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int row = row0 + r0 + hf * 8;
      if (row < n_rows) *reinterpret_cast<uint32_t*>(dst + (size_t)row * ld + t * 8 + c0) = pack_bf16x2(o[t][2 * hf] * sc, o[t][2 * hf + 1] * sc);
    }
}

__global__ void __launch_bounds__(DA_THREADS, 1) dec_self_attn_fwd_warp_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int n = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= p.heads) return;
  bf16* sK = reinterpret_cast<bf16*>(da_smem + (size_t)h * SW_FWD_BYTES);
  bf16* sV = sK + 32 * MM_LDK;
  bf16* sQ = sV + 32 * MM_LDK;
  bf16* sP = sQ + 32 * MM_LDK;
  stage_rows_warp(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, p.ldk, lane);
  stage_rows_warp(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, p.ldv, lane);
  stage_rows_warp(sQ, p.q + (size_t)n * p.tq * p.ldq + h * DA_D, p.tq, p.ldq, lane);
  stage_commit_wait();
  __syncwarp();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int r0 = lane >> 2, c0 = 2 * (lane & 3);
  const unsigned long long bh = (unsigned long long)n * p.heads + h;
  float acc[2][4][4];
  qk_block(acc, sQ, sK, 0, lane);
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = mt * 16 + r0 + (e >> 1) * 8, key = nt * 8 + c0 + (e & 1);
        const float sc = key_visible(p, n, row, key) ? acc[mt][nt][e] * p.scale : -INFINITY;
        acc[mt][nt][e] = sc;
        mx[mt * 2 + (e >> 1)] = fmaxf(mx[mt * 2 + (e >> 1)], sc);
      }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = mt * 2 + (e >> 1);
        const float sc = acc[mt][nt][e];
        const float pr = (sc == -INFINITY) ? 0.f : __expf(sc - mx[i]);
        acc[mt][nt][e] = pr;
        sum[i] += pr;
      }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
    const int row = r0 + 8 * i;
    if ((lane & 3) == 0 && row < p.tq && p.lse != nullptr) p.lse[(bh)*p.tq + row] = mx[i] + __logf(sum[i]);
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int i = mt * 2 + hf, row = mt * 16 + r0 + hf * 8, key = nt * 8 + c0;
        const float inv = sum[i] > 0.f ? 1.0f / sum[i] : 0.f;
        float p0 = acc[mt][nt][2 * hf] * inv, p1 = acc[mt][nt][2 * hf + 1] * inv;
        if (p.p_drop > 0.f) {
          const unsigned long long idx = ((bh * p.tq + row) << 8) + key;
          p0 = keep_elem(p.seed, idx, thresh) ? p0 * inv_keep : 0.f;
          p1 = keep_elem(p.seed, idx + 1, thresh) ? p1 * inv_keep : 0.f;
        }
        *reinterpret_cast<uint32_t*>(sP + row * MM_LDS + key) = pack_bf16x2(p0, p1);
      }
  __syncwarp();
  bf16* obase = p.o + (size_t)n * p.tq * p.ldo + h * DA_D;
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    float o[8][4];
    small_mm<false>(o, sP, mt * 16, sV, lane);
    store_tile_rows(obase, p.ldo, mt * 16, p.tq, o, 1.0f, lane);
  }
}

__global__ void __launch_bounds__(DA_THREADS, 1) dec_self_attn_bwd_warp_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int n = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= p.heads) return;
  bf16* sK = reinterpret_cast<bf16*>(da_smem + (size_t)h * SW_BWD_BYTES);
  bf16* sV = sK + 32 * MM_LDK;
  bf16* sQ = sV + 32 * MM_LDK;
  bf16* sDO = sQ + 32 * MM_LDK;
  bf16* sP = sDO + 32 * MM_LDK;
  bf16* sDS = sP + 32 * MM_LDS;
  float* sLse = reinterpret_cast<float*>(sDS + 32 * MM_LDS);
  float* sDelta = sLse + 32;
  stage_rows_warp(sK, p.k + (size_t)n * p.tk * p.ldk + h * DA_D, p.tk, p.ldk, lane);
  stage_rows_warp(sV, p.v + (size_t)n * p.tk * p.ldv + h * DA_D, p.tk, p.ldv, lane);
  stage_rows_warp(sQ, p.q + (size_t)n * p.tq * p.ldq + h * DA_D, p.tq, p.ldq, lane);
  stage_rows_warp(sDO, p.d_o + (size_t)n * p.tq * p.ldo + h * DA_D, p.tq, p.ldo, lane);
  const unsigned long long bh = (unsigned long long)n * p.heads + h;
  {                                                               // lane = query row: delta_i = dO_i . O_i, lse_i
    float d = 0.f, l = 0.f;
    if (lane < p.tq) {
      const size_t off = ((size_t)n * p.tq + lane) * p.ldo + h * DA_D;
      const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + off);
      const uint4* po = reinterpret_cast<const uint4*>(p.o + off);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = __ldg(pd + c), b = __ldg(po + c);
        d += bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y) +
             bf16lo(a.z) * bf16lo(b.z) + bf16hi(a.z) * bf16hi(b.z) + bf16lo(a.w) * bf16lo(b.w) + bf16hi(a.w) * bf16hi(b.w);
      }
      l = p.lse[bh * p.tq + lane];
    }
    sDelta[lane] = d;
    sLse[lane] = l;
  }
  stage_commit_wait();
  __syncwarp();
  const uint32_t thresh = drop_threshold(p.p_drop);
  const float inv_keep = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const int r0 = lane >> 2, c0 = 2 * (lane & 3);
  {
    float acc[2][4][4], pr[2][4][4];
    qk_block(acc, sQ, sK, 0, lane);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = mt * 16 + r0 + (e >> 1) * 8, key = nt * 8 + c0 + (e & 1);
          pr[mt][nt][e] = key_visible(p, n, row, key) ? __expf(acc[mt][nt][e] * p.scale - sLse[row]) : 0.f;
        }
    qk_block(acc, sDO, sV, 0, lane);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int row = mt * 16 + r0 + hf * 8, key = nt * 8 + c0;
          float pt[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pe = pr[mt][nt][2 * hf + e];
            float dp = acc[mt][nt][2 * hf + e];
            pt[e] = pe;
            if (p.p_drop > 0.f) {
              const unsigned long long idx = ((bh * p.tq + row) << 8) + key + e;
              const bool kp = keep_elem(p.seed, idx, thresh);
              pt[e] = kp ? pe * inv_keep : 0.f;
              dp = kp ? dp * inv_keep : 0.f;
            }
            ds[e] = pe * (dp - sDelta[row]);
          }
          *reinterpret_cast<uint32_t*>(sP + row * MM_LDS + key) = pack_bf16x2(pt[0], pt[1]);
          *reinterpret_cast<uint32_t*>(sDS + row * MM_LDS + key) = pack_bf16x2(ds[0], ds[1]);
        }
  }
  __syncwarp();
  float o[8][4];
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {                                // dQ = scale * dS K
    small_mm<false>(o, sDS, mt * 16, sK, lane);
    store_tile_rows(p.dq + (size_t)n * p.tq * p.lddq + h * DA_D, p.lddq, mt * 16, p.tq, o, p.scale, lane);
  }
#pragma unroll 1
  for (int kt = 0; kt < 2; ++kt) {                                // dV = P~^T dO ; dK = scale * dS^T Q
    if (kt * 16 >= p.tk) break;
    small_mm<true>(o, sP, kt * 16, sDO, lane);
    store_tile_rows(p.dv + (size_t)n * p.tk * p.lddv + h * DA_D, p.lddv, kt * 16, p.tk, o, 1.0f, lane);
    small_mm<true>(o, sDS, kt * 16, sQ, lane);
    store_tile_rows(p.dk + (size_t)n * p.tk * p.lddk + h * DA_D, p.lddk, kt * 16, p.tk, o, p.scale, lane);
  }
}

static size_t dec_attn_mma_smem(int tk, bool bwd) {
  const int tkp = (tk + 31) & ~31;
  if (!bwd) return (size_t)(2 * tkp * MM_LDK + 32 * MM_LDK + 32 * MM_LDP) * 2 + (size_t)(DA_WARPS * 32 + 32) * 4;
  const int v_elems = tkp * MM_LDK > 2 * 32 * MM_LDP ? tkp * MM_LDK : 2 * 32 * MM_LDP;
  return (size_t)(tkp * MM_LDK + v_elems + 2 * 32 * MM_LDK) * 2 + 64 * 4;
}

static int g_dec_attn_variant = 1;          // 1 = mma.sync tensor-core kernels (default), 0 = scalar CUDA-core kernels

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
  static bool attr = false;
  if (!attr) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_smem(DA_MAX_TQ, DA_MAX_TK, false)));
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_mma_smem(DA_MAX_TK, false)));
    attr = true;
  }
  static bool attr_w = false;
  if (!attr_w) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_self_attn_fwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_WARPS * SW_FWD_BYTES));
    attr_w = true;
  }
  if (g_dec_attn_variant == 1 && tk <= 32 && tq <= 32 && heads <= DA_WARPS)
    dec_self_attn_fwd_warp_kernel<<<n, DA_THREADS, DA_WARPS * SW_FWD_BYTES, stream>>>(p);
  else if (g_dec_attn_variant == 1) dec_attn_fwd_mma_kernel<<<n * heads, DA_THREADS, dec_attn_mma_smem(tk, false), stream>>>(p);
  else dec_attn_fwd_kernel<<<n * heads, DA_THREADS, dec_attn_smem(tq, tk, false), stream>>>(p);
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
  static bool attr = false;
  if (!attr) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_smem(DA_MAX_TQ, DA_MAX_TK, true)));
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_attn_mma_smem(DA_MAX_TK, true)));
    attr = true;
  }
  static bool attr_w = false;
  if (!attr_w) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(dec_self_attn_bwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_WARPS * SW_BWD_BYTES));
    attr_w = true;
  }
  if (g_dec_attn_variant == 1 && tk <= 32 && tq <= 32 && heads <= DA_WARPS)
    dec_self_attn_bwd_warp_kernel<<<n, DA_THREADS, DA_WARPS * SW_BWD_BYTES, stream>>>(p);
  else if (g_dec_attn_variant == 1) dec_attn_bwd_mma_kernel<<<n * heads, DA_THREADS, dec_attn_mma_smem(tk, true), stream>>>(p);
  else dec_attn_bwd_kernel<<<n * heads, DA_THREADS, dec_attn_smem(tq, tk, true), stream>>>(p);
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

// A/B switch: 1 = mma.sync tensor-core decoder attention (default), 0 = scalar CUDA-core kernels
extern "C" int ccd_set_dec_attn_variant(int v) {
  g_dec_attn_variant = v ? 1 : 0;
  return CCD_OK;
}
