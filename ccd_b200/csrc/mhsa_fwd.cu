// Fused multi-head self-attention forward on tcgen05/TMEM for the CCD ViT encoder
// (reference: Attention.forward, Dino/modules/vision_transformer.py:80-92 -- softmax(q k^T * hd^-0.5) v over
//  N = 256 tokens, head_dim 64, no mask, no dropout).  The reference materialises [2B,H,256,256] fp32 scores;
//  here scores never leave the SM.
//
// One CTA per (sequence, head, 128-query tile); two CTAs co-resident per SM (80 KB smem, 256 TMEM columns each) so
// one CTA's softmax overlaps the other's MMAs.
//   warp 0   : TMA producer -- Q tile [128x64], K [256x64], V [256x64] straight out of the fused qkv activation
//              [T, 3E] (no head-split copies), SWIZZLE_128B
//   warp 1   : TMEM alloc + UMMA issuer:  S[128x256] = Q K^T  (4 x tcgen05.mma 128x256x16, fp32 in TMEM cols 0..255)
//                                         O[128x64]  = P V    (16 x tcgen05.mma 128x64x16; V read MN-major)
//   warps 2-5: one query row per thread: row max, exp2 with folded scale, bf16 P written back over S
//              (TMEM cols 0..127, A-operand of the second MMA; P_IN_TMEM) or into smem (fallback layout),
//              then O / rowsum -> bf16 out[T,E], and the per-row log2-sum-exp for backward.
#include "ccd_common.cuh"
#include "tmap.cuh"

namespace ccd {

constexpr int ATT_N = 256;    // tokens per sequence (8x32 grid, no CLS: vision_transformer.py:229-238)
constexpr int ATT_D = 64;     // head dim (all three archs)
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_THREADS = 192;

struct MhsaFwdParams {
  bf16* out;        // [T, E]
  float* lse2;      // [S, H, 256]  log2-domain: max*c + log2(sum),  c = scale*log2(e)
  int E, H;
  float scale_log2;  // hd^-0.5 * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool P_IN_TMEM>
struct MhsaFwdCfg {
  static constexpr int Q_BYTES = ATT_BM * ATT_D * 2;   // 16 KB
  static constexpr int KV_BYTES = ATT_N * ATT_D * 2;   // 32 KB
  static constexpr int P_BYTES = P_IN_TMEM ? 0 : ATT_BM * ATT_N * 2;  // 64 KB
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
};

template <bool P_IN_TMEM>
__global__ void __launch_bounds__(ATT_THREADS, P_IN_TMEM ? 2 : 1)
mhsa_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const MhsaFwdParams p) {
  using Cfg = MhsaFwdCfg<P_IN_TMEM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mt = blockIdx.x;   // query tile 0/1
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int row0 = s * ATT_N;  // first token row of this sequence in [T, 3E]

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;          // S: cols [0,256)
  const uint32_t tP = tmem_base;          // P (bf16 pairs) aliases cols [0,128)
  const uint32_t tO = tmem_base + 128;    // O: cols [128,192) -- free once S has been consumed

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_qk, Cfg::Q_BYTES + Cfg::KV_BYTES);
      tma_load_2d(sQ, &tmQKV, bar_qk, h * ATT_D, row0 + mt * ATT_BM);
      tma_load_2d(sK, &tmQKV, bar_qk, p.E + h * ATT_D, row0);
      tma_load_2d(sK + 16384, &tmQKV, bar_qk, p.E + h * ATT_D, row0 + 128);
      mbar_arrive_expect_tx(bar_v, Cfg::KV_BYTES);
      tma_load_2d(sV, &tmQKV, bar_v, 2 * p.E + h * ATT_D, row0);
      tma_load_2d(sV + 16384, &tmQKV, bar_v, 2 * p.E + h * ATT_D, row0 + 128);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- S = Q K^T : M=128, N=256, K=64 (4 k-steps), both operands K-major ----
      const uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)
        umma_ss(tS, umma_smem_desc_sw128(q_addr + k * 32, 16, 1024), umma_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
      // ---- O = P V : M=128, N=64, K=256 (16 k-steps); V[token, hd] is the MN-major B operand ----
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(sV), p_addr = smem_u32(sP);
#pragma unroll
      for (int k = 0; k < ATT_N / 16; ++k) {
        const uint64_t dv = umma_smem_desc_sw128(v_addr + k * 2048, 8192, 1024);
        if constexpr (P_IN_TMEM) {
          umma_ts(tO, tP + k * 8, dv, idesc_o, k > 0 ? 1u : 0u);  // 16 bf16 = 8 TMEM columns per k-step
        } else {
          const uint64_t dp = umma_smem_desc_sw128(p_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          umma_ss(tO, dp, dv, idesc_o, k > 0 ? 1u : 0u);
        }
      }
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ---- softmax + epilogue: thread <-> query row (TMEM lane) ----
    const int q = warp & 3;
    const int r = q * 32 + lane;                         // row within the tile
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = p.scale_log2;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < ATT_N / 32; ++ch) {
      uint32_t raw[32];
      tmem_ld_32x32(tS + lane_sel + ch * 32, raw);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
    }
    const float mc = mx * c;
    float sum = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < ATT_N / 32; ++ch) {
      uint32_t raw[32];
      tmem_ld_32x32(tS + lane_sel + ch * 32, raw);
      tmem_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(raw[2 * j]), c, -mc));
        const float p1 = ex2_approx(fmaf(__uint_as_float(raw[2 * j + 1]), c, -mc));
        pk[j] = pack_bf16x2(p0, p1);
        sum += bf16lo(pk[j]) + bf16hi(pk[j]);   // sum the rounded values: P V / sum(P) stays a convex combination
      }
      if constexpr (P_IN_TMEM) {
        tmem_st_32x16(tP + lane_sel + ch * 16, pk);
      } else {
        // K-major SWIZZLE_128B tile: 64-key block kb at kb*16 KB, row r at r*128 B, 16-byte chunk index ^ (r & 7)
        uint8_t* rowp = sP + (ch >> 1) * 16384 + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int chunk = ((ch & 1) * 4 + v4) ^ (r & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[4 * v4], pk[4 * v4 + 1], pk[4 * v4 + 2], pk[4 * v4 + 3]);
        }
      }
    }
    if constexpr (P_IN_TMEM) {
      tmem_wait_st();
      tc_fence_before();
    } else {
      fence_proxy_async_smem();
    }
    mbar_arrive(bar_p);

    // ---- epilogue ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const size_t tok = (size_t)row0 + mt * ATT_BM + r;
    bf16* orow = p.out + tok * p.E + h * ATT_D;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t raw[32];
      tmem_ld_32x32(tO + lane_sel + half * 32, raw);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(raw[8 * g + 0]) * inv, __uint_as_float(raw[8 * g + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(raw[8 * g + 2]) * inv, __uint_as_float(raw[8 * g + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(raw[8 * g + 4]) * inv, __uint_as_float(raw[8 * g + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(raw[8 * g + 6]) * inv, __uint_as_float(raw[8 * g + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + half * 32 + g * 8) = o;
      }
    }
    if (p.lse2 != nullptr) p.lse2[((size_t)s * p.H + h) * ATT_N + mt * ATT_BM + r] = mc + log2f(sum);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <bool P_IN_TMEM>
static int launch_mhsa_fwd(const CUtensorMap& tm, const MhsaFwdParams& p, int S, cudaStream_t stream) {
  using Cfg = MhsaFwdCfg<P_IN_TMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(mhsa_fwd_kernel<P_IN_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(ATT_N / ATT_BM, p.H, S);
  mhsa_fwd_kernel<P_IN_TMEM><<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, p);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Persistent variant (variant 2): one CTA per SM loops over (sequence, head) work items; both 128-query tiles of an
// item share ONE K/V load; Q/K/V shared-memory buffers are double buffered (TMA prefetch of item i+1 under item i),
// two softmax warpgroups (one per query tile) ping-pong on the MUFU while the single MMA thread alternates
// S_t = Q_t K^T and O_t = P_t V.  TMEM: S0 [0,256) S1 [256,512); P_t aliases the first 128 columns of S_t, O_t the
// next 64.  320 threads: warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 softmax tile 0, warps 6-9 softmax tile 1.
// ---------------------------------------------------------------------------------------------------------
constexpr int ATT2_THREADS = 320;
constexpr int ATT2_BUF = 3 * ATT_N * ATT_D * 2;        // Q + K + V = 96 KB
constexpr int ATT2_STAGE = ATT_BM * ATT_D * 2;         // 16 KB: one [128 x 64] bf16 output tile per softmax warpgroup (TMA store)
constexpr int ATT2_SMEM = 2 * ATT2_BUF + 2 * ATT2_STAGE + 256 + 1024;

__global__ void __launch_bounds__(ATT2_THREADS, 1)
mhsa_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const MhsaFwdParams p,
                           int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * ATT2_BUF + 2 * ATT2_STAGE);
  uint64_t* full_qk = bars;        // [2]
  uint64_t* full_v = bars + 2;     // [2]
  uint64_t* empty = bars + 4;      // [2]
  uint64_t* s_full = bars + 6;     // [2] per query tile
  uint64_t* p_full = bars + 8;     // [2]
  uint64_t* o_full = bars + 10;    // [2]
  uint64_t* tmem_free = bars + 12; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_qk[i], 1);
      mbar_init(&full_v[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&tmem_free[i], 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  pdl_launch_dependents();
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_full_base(tmem_slot);      // all 512 columns: literal 0 (see ccd_common.cuh)
  pdl_wait();                                      // qkv of the preceding GEMM is complete and visible

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t kph = (it >> 1) & 1;
        const int s = w / p.H, h = w - s * p.H;
        const int row0 = s * ATT_N;
        uint8_t* buf = smem + b * ATT2_BUF;
        mbar_wait(&empty[b], kph ^ 1);
        mbar_arrive_expect_tx(&full_qk[b], 2 * ATT_N * ATT_D * 2);
        tma_load_2d(buf, &tmQKV, &full_qk[b], h * ATT_D, row0);
        tma_load_2d(buf + 16384, &tmQKV, &full_qk[b], h * ATT_D, row0 + 128);
        tma_load_2d(buf + 32768, &tmQKV, &full_qk[b], p.E + h * ATT_D, row0);
        tma_load_2d(buf + 49152, &tmQKV, &full_qk[b], p.E + h * ATT_D, row0 + 128);
        mbar_arrive_expect_tx(&full_v[b], ATT_N * ATT_D * 2);
        tma_load_2d(buf + 65536, &tmQKV, &full_v[b], 2 * p.E + h * ATT_D, row0);
        tma_load_2d(buf + 81920, &tmQKV, &full_v[b], 2 * p.E + h * ATT_D, row0 + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      // Issue order (software pipelined across items so that the two softmax warpgroups run half an item out of phase and
      // the MUFU never idles while a tile waits for its MMAs):   S0(0) | S1(i)  O0(i)  S0(i+1)  O1(i) | ...
      // S_t(i+1) overwrites the TMEM columns of P_t(i) / O_t(i): it is issued after O_t(i) (MMAs retire in issue order)
      // and after the epilogue of tile t has read O_t(i) out (tmem_free[t]).
      // one base descriptor per layout; each MMA advances it by a constant (the 16 N = 64 MMAs of O = P V execute in 32 cycles
      // each: deriving descriptors per MMA on the single issuing thread cost more than the MMAs themselves)
      const uint64_t d_k = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);          // K-major tiles (Q, K)
      const uint64_t d_v = umma_smem_desc_sw128(smem_u32(smem), 8192, 1024);        // V read MN-major
      auto issue_s = [&](int t, uint32_t buf_off) {
        const uint64_t a = umma_desc_advance(d_k, buf_off + t * 16384), b = umma_desc_advance(d_k, buf_off + 32768);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_ss(tmem_base + t * 256, umma_desc_advance(a, k * 32), umma_desc_advance(b, k * 32), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      auto issue_o = [&](int t, uint32_t buf_off) {
        const uint64_t b = umma_desc_advance(d_v, buf_off + 65536);
#pragma unroll
        for (int k = 0; k < ATT_N / 16; ++k)
          umma_ts(tmem_base + t * 256 + 128, tmem_base + t * 256 + k * 8, umma_desc_advance(b, k * 2048), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
      };
      // (A polling issuer that serves the two tiles as independent chains was measured 6-12 % SLOWER than this fixed
      //  order: its try_wait loop competes for issue slots with the two softmax warps of its scheduler.)
      int it = 0;
      if (blockIdx.x < n_items) {
        mbar_wait(&full_qk[0], 0);
        tc_fence_after();
        issue_s(0, 0u);
      }
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t kph = (it >> 1) & 1, ph = it & 1;
        const uint32_t buf = (uint32_t)(b * ATT2_BUF);          // byte offset of this item's Q/K/V buffer
        mbar_wait(&tmem_free[1], ph ^ 1);         // previous item's O_1 has been read out of TMEM
        tc_fence_after();
        issue_s(1, buf);
        mbar_wait(&full_v[b], kph);
        mbar_wait(&p_full[0], ph);
        tc_fence_after();
        issue_o(0, buf);
        if (w + (int)gridDim.x < n_items) {       // S_0 of the next item (other Q/K/V buffer)
          const int it1 = it + 1;
          mbar_wait(&full_qk[it1 & 1], (it1 >> 1) & 1);
          mbar_wait(&tmem_free[0], (it1 & 1) ^ 1);
          tc_fence_after();
          issue_s(0, (uint32_t)((it1 & 1) * ATT2_BUF));
        }
        mbar_wait(&p_full[1], ph);
        tc_fence_after();
        issue_o(1, buf);
        umma_commit(&empty[b]);                   // Q/K/V buffer reusable once every MMA of this item retired
      }
    }
    __syncwarp();
  } else {
    const int t = (warp - 2) >> 2;                // query tile / softmax warpgroup
    const int q = warp & 3;                       // TMEM lane quarter
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + t * 256, tO = tS + 128;
    const float c = p.scale_log2;
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int s = w / p.H, h = w - s * p.H;
      mbar_wait(&s_full[t], ph);
      tc_fence_after();
      // two 32-column TMEM loads in flight per wait (64-column chunks) in both passes
      float mx = -INFINITY;
#pragma unroll 1
      for (int ch = 0; ch < ATT_N / 64; ++ch) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tS + lane_sel + ch * 64, ra);
        tmem_ld_32x32(tS + lane_sel + ch * 64 + 32, rb);
        tmem_wait_ld();
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          m0 = fmaxf(m0, __uint_as_float(ra[j]));
          m1 = fmaxf(m1, __uint_as_float(rb[j]));
        }
        mx = fmaxf(mx, fmaxf(m0, m1));
      }
      const float mc = mx * c;
      // packed fp32x2 scale and sum (FFMA2 / FADD2): 5 issue slots per pair of scores instead of 9.  The normaliser sums
      // the unrounded probabilities (what the saved log-sum-exp must describe for the backward's recomputation of P).
      const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
      float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int ch = 0; ch < ATT_N / 64; ++ch) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tS + lane_sel + ch * 64, ra);
        tmem_ld_32x32(tS + lane_sel + ch * 64 + 32, rb);
        tmem_wait_ld();
        uint32_t pa[16], pb[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 xa = __ffma2_rn(make_float2(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1])), c2, nmc2);
          const float2 xb = __ffma2_rn(make_float2(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1])), c2, nmc2);
          const float2 ea = make_float2(ex2_approx(xa.x), ex2_approx(xa.y));
          const float2 eb = make_float2(ex2_approx(xb.x), ex2_approx(xb.y));
          pa[j] = pack_bf16x2(ea.x, ea.y);
          pb[j] = pack_bf16x2(eb.x, eb.y);
          sa = __fadd2_rn(sa, ea);
          sb = __fadd2_rn(sb, eb);
        }
        tmem_st_32x16(tS + lane_sel + ch * 32, pa);        // P chunk (bf16 pairs) over S columns already consumed
        tmem_st_32x16(tS + lane_sel + ch * 32 + 16, pb);
      }
      const float sum0 = sa.x + sa.y, sum1 = sb.x + sb.y;
      const float sum = sum0 + sum1;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t]);

      mbar_wait(&o_full[t], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      // O leaves TMEM first and the tile is handed back (S_t of the next item may start); the normalised bf16 rows then go to this
      // warpgroup's 16 KB staging tile in the SWIZZLE_128B layout and leave with ONE bulk tensor store (the thread-per-row form,
      // 8 x 16-byte global stores per thread to 32 different 128-byte lines per warp instruction, serialised in the LSU).
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(tO + lane_sel, o0);
      tmem_ld_32x32(tO + lane_sel + 32, o1);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&tmem_free[t]);
      const bool elected = (q == ((2 + 4 * t) & 3)) && lane == 0;     // lane 0 of the warpgroup's first warp
      if (elected) tma_store_wait_read();                 // the previous item's store has finished reading the staging tile
      named_bar_sync(1 + t, 128);
      const uint32_t srow = smem_u32(smem + 2 * ATT2_BUF + t * ATT2_STAGE) + (uint32_t)r * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((g ^ (r & 7)) * 16)),
                     "r"(pack_bf16x2(__uint_as_float(o0[8 * g + 0]) * inv, __uint_as_float(o0[8 * g + 1]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o0[8 * g + 2]) * inv, __uint_as_float(o0[8 * g + 3]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o0[8 * g + 4]) * inv, __uint_as_float(o0[8 * g + 5]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o0[8 * g + 6]) * inv, __uint_as_float(o0[8 * g + 7]) * inv))
                     : "memory");
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)(((4 + g) ^ (r & 7)) * 16)),
                     "r"(pack_bf16x2(__uint_as_float(o1[8 * g + 0]) * inv, __uint_as_float(o1[8 * g + 1]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o1[8 * g + 2]) * inv, __uint_as_float(o1[8 * g + 3]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o1[8 * g + 4]) * inv, __uint_as_float(o1[8 * g + 5]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(o1[8 * g + 6]) * inv, __uint_as_float(o1[8 * g + 7]) * inv))
                     : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (elected) {
        tma_store_2d(&tmO, smem + 2 * ATT2_BUF + t * ATT2_STAGE, h * ATT_D, s * ATT_N + t * ATT_BM);
        tma_store_commit();
      }
      if (p.lse2 != nullptr) p.lse2[((size_t)s * p.H + h) * ATT_N + t * ATT_BM + r] = mc + log2f(sum);
    }
  }

  if (warp >= 2 && (warp & 3) == 2 && lane == 0) tma_store_wait_read();   // the elected threads: staging tiles drained before exit
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int launch_mhsa_fwd_persistent(const CUtensorMap& tm, const CUtensorMap& tmO, const MhsaFwdParams& p, int S, cudaStream_t stream) {
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(mhsa_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
    int dev = 0;
    CCD_CUDA_CHECK(cudaGetDevice(&dev));
    CCD_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int n_items = S * p.H;
  const int grid = n_items < num_sms ? n_items : num_sms;
  CCD_CUDA_CHECK(launch_pdl(mhsa_fwd_persistent_kernel, dim3(grid), dim3(ATT2_THREADS), (size_t)ATT2_SMEM, stream, tm, tmO, p, n_items));
  return CCD_OK;
}

}  // namespace ccd

using namespace ccd;

// C ABI -- see include/ccd_b200.h
extern "C" int ccd_mhsa_fwd(const void* qkv, void* out, float* lse2, int S, int H, int variant, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!qkv || !out || S <= 0 || H <= 0) return CCD_ERR_ARG;
  const int E = H * ATT_D;
  CUtensorMap tm;
  if (!get_tmap_bf16_2d(&tm, qkv, (uint64_t)S * ATT_N, (uint64_t)3 * E, (uint64_t)3 * E, 128, 64)) return CCD_ERR_TMAP;
  MhsaFwdParams p;
  p.out = reinterpret_cast<bf16*>(out);
  p.lse2 = lse2;
  p.E = E;
  p.H = H;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  if (variant == 1) return launch_mhsa_fwd<false>(tm, p, S, stream);
  if (variant == 2) {
    CUtensorMap tmO;                                       // output tiles leave through TMA stores
    if (!get_tmap_bf16_2d(&tmO, out, (uint64_t)S * ATT_N, (uint64_t)E, (uint64_t)E, 128, 64)) return CCD_ERR_TMAP;
    return launch_mhsa_fwd_persistent(tm, tmO, p, S, stream);
  }
  return launch_mhsa_fwd<true>(tm, p, S, stream);
}
