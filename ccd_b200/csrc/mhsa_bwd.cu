// Fused multi-head self-attention backward on tcgen05/TMEM (backward of Attention.forward,
// Dino/modules/vision_transformer.py:80-92; the reference gets it from autograd over the materialised
// [2B,H,256,256] probabilities, here P is recomputed from the saved per-row log-sum-exp).
//
// One CTA per (sequence, head); Q,K,V,dO [256x64] bf16 resident in shared memory (TMA, SWIZZLE_128B), all five
// contractions on tcgen05.mma with fp32 accumulators in TMEM (all 512 columns):
//   for key tile j (128 keys), query tile i (128 queries):
//     S^T  = K_j Q_i^T            [keys x queries]   cols [0,128)
//     dP^T = V_j dO_i^T           [keys x queries]   cols [128,256)
//     P^T  = exp2(S^T c - lse2[q]);  dS^T = P^T (dP^T - delta[q]) * scale     (thread = key row)
//     dV_j += P^T  dO_i           cols [320,384)     (dO read MN-major)
//     dK_j += dS^T Q_i            cols [256,320)     (Q  read MN-major; pipelined kernel: dS^T as a TMEM A operand)
//     dQ_i += dS   K_j            cols [384,512)     (dS^T tile read MN-major as A, K read MN-major)
// P^T / dS^T go through shared memory as bf16 in the 128B-swizzled layout that is simultaneously a K-major A tile
// (for dV, dK) and an MN-major A tile (for dQ).
#include "ccd_common.cuh"
#include "tmap.cuh"

namespace ccd {

constexpr int ATB_N = 256;
constexpr int ATB_D = 64;
constexpr int ATB_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 = two softmax-gradient warpgroups

struct MhsaBwdParams {
  const bf16* o;      // [T, E] forward output
  const bf16* d_o;    // [T, E]
  const float* lse2;  // [S, H, 256]
  const float* delta; // [S, H, 256]  rowsum(O * dO), produced by mhsa_delta_kernel (pipelined kernel: -delta * scale)
  const float* nlse;  // [S, H, 256]  -lse2 (pipelined kernel only)
  bf16* dqkv;         // [T, 3E]
  float* dbias;       // optional [3E] (caller zero-fills): column sums of dqkv = gradient of Attention.qkv.bias
  int E, H;
  float scale, scale_log2;
};

struct MhsaBwdSmem {
  static constexpr int TILE = ATB_N * ATB_D * 2;  // 32 KB
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = TILE;
  static constexpr int OFF_V = 2 * TILE;
  static constexpr int OFF_DO = 3 * TILE;
  static constexpr int OFF_PT = 4 * TILE;   // [128 keys x 128 queries] bf16, two 64-query blocks of 16 KB
  static constexpr int OFF_DST = 5 * TILE;
  static constexpr int OFF_LSE = 6 * TILE;  // 256 f32
  static constexpr int OFF_DELTA = OFF_LSE + 1024;
  static constexpr int OFF_BAR = OFF_DELTA + 1024;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
};

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  return bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y) +
         bf16lo(a.z) * bf16lo(b.z) + bf16hi(a.z) * bf16hi(b.z) + bf16lo(a.w) * bf16lo(b.w) + bf16hi(a.w) * bf16hi(b.w);
}

// 64 fp32 TMEM columns of this thread's lane -> bf16 -> 128 contiguous bytes in global memory
__device__ __forceinline__ void store_tmem_row64(uint32_t taddr, bf16* dst) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t raw[32];
    tmem_ld_32x32(taddr + half * 32, raw);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(raw[8 * g + 0]), __uint_as_float(raw[8 * g + 1]));
      o.y = pack_bf16x2(__uint_as_float(raw[8 * g + 2]), __uint_as_float(raw[8 * g + 3]));
      o.z = pack_bf16x2(__uint_as_float(raw[8 * g + 4]), __uint_as_float(raw[8 * g + 5]));
      o.w = pack_bf16x2(__uint_as_float(raw[8 * g + 6]), __uint_as_float(raw[8 * g + 7]));
      *reinterpret_cast<uint4*>(dst + half * 32 + g * 8) = o;
    }
  }
}

// delta[s,h,q] = sum_d O[q, h*64+d] * dO[q, h*64+d]: one warp per token row, 16-byte loads, 8 lanes per head.
// (Inside the main kernel this prologue was ~30 % of the stall samples: 32 rows x 16 B per load instruction.)
// nlse != NULL (pipelined main kernel): writes -delta * scale and nlse = -lse2, the forms its inner loop consumes, so that the
// TMA producer can drop both vectors into shared memory with two bulk copies per item.
__global__ void __launch_bounds__(256) mhsa_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                         float* __restrict__ delta, const float* __restrict__ lse2,
                                                         float* __restrict__ nlse, float scale, int T, int E, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const uint4* po = reinterpret_cast<const uint4*>(o + (size_t)row * E);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + (size_t)row * E);
  const int nchunk = E >> 3;                       // 8 bf16 per 16-byte chunk; 8 chunks per head
  const int s = row >> 8, q = row & 255;
  for (int c0 = 0; c0 < nchunk; c0 += 32) {
    const int c = c0 + lane;
    float acc = (c < nchunk) ? dot8_bf16(po[c], pd[c]) : 0.f;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((lane & 7) == 0 && c < nchunk) {
      const size_t idx = ((size_t)s * H + (c >> 3)) * ATB_N + q;
      if (nlse != nullptr) {
        delta[idx] = -acc * scale;
        nlse[idx] = -lse2[idx];
      } else {
        delta[idx] = acc;
      }
    }
  }
}

// qkv-bias gradient = column sums of dqkv:
//   query part  sum_q dQ[q, d]: produced by the pipelined kernel itself from the TMEM tiles as they are written out (below);
//   value part  sum_keys dV[key, d] = sum_q (sum_keys P[q, key]) dO[q, d] = sum_q dO[q, d] (softmax rows sum to 1) = the column
//               sums of d_o = (column sums of the proj-layer output gradient) . W_proj: a tiny vector-matrix product the caller
//               does from the proj-bias gradient it already holds (ops.mhsa_bwd / ccd_vecmat_add_f32) -- no reduction here;
//   key part    sum_keys dK[key, d] = sum_q (sum_keys dS[q, key]) Q[q, d] = 0 identically, because sum_keys dS[q, :] =
//   scale (P . dP - delta[q]) = 0 (a key bias shifts every score of a row by the same amount): left untouched.
// Per-CTA partial sums live in shared memory across all items of the persistent CTA and are flushed with one global atomic
// per (head, column) at the end (per-item global atomics concentrate on ~12 L2 slices and throttle: measured 0.65 ms).
// 64 fp32 TMEM columns of this thread's lane -> bf16 -> 128 contiguous bytes in global memory, and (colsum != NULL) the sums
// of the 64 columns over the 32 rows of this warp -> atomicAdd(colsum[0..63]) (shared memory).  Value-halving butterfly:
// 31 shuffles per 32 columns, lane l ends up with the total of column l.
__device__ __forceinline__ void store_tmem_row64_colsum(uint32_t taddr, bf16* dst, float* colsum, int lane) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t raw[32];
    tmem_ld_32x32(taddr + half * 32, raw);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(raw[8 * g + 0]), __uint_as_float(raw[8 * g + 1]));
      o.y = pack_bf16x2(__uint_as_float(raw[8 * g + 2]), __uint_as_float(raw[8 * g + 3]));
      o.z = pack_bf16x2(__uint_as_float(raw[8 * g + 4]), __uint_as_float(raw[8 * g + 5]));
      o.w = pack_bf16x2(__uint_as_float(raw[8 * g + 6]), __uint_as_float(raw[8 * g + 7]));
      *reinterpret_cast<uint4*>(dst + half * 32 + g * 8) = o;
    }
    if (colsum != nullptr) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(raw[k]);
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int k = 0; k < step; ++k) {
          const float keep = up ? v[k + step] : v[k];
          const float send = up ? v[k] : v[k + step];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
      }
      atomicAdd(colsum + half * 32 + lane, v[0]);
    }
  }
}

// 64 fp32 TMEM columns of this thread's lane -> bf16 -> this thread's 128-byte row of a [128 x 64] shared tile in the
// SWIZZLE_128B layout a TMA store expects (chunk c of row r at c ^ (r & 7)): conflict-free 16-byte stores, and the global
// write becomes ONE bulk tensor store per tile instead of 8 x 128 scattered 16-byte stores per warp (each warp-level STG.128 of
// the thread-per-row form touches 32 different 128-byte lines: the LSU serialises them, ~2400 cycles per tile measured).
// colsum != NULL: also the column sums over the warp's 32 rows (value-halving butterfly, see store_tmem_row64_colsum).
__device__ __forceinline__ void stage_tmem_row64(uint32_t taddr, uint32_t srow, int r, float* colsum, int lane) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t raw[32];
    tmem_ld_32x32(taddr + half * 32, raw);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t a = pack_bf16x2(__uint_as_float(raw[8 * g + 0]), __uint_as_float(raw[8 * g + 1]));
      const uint32_t b = pack_bf16x2(__uint_as_float(raw[8 * g + 2]), __uint_as_float(raw[8 * g + 3]));
      const uint32_t c = pack_bf16x2(__uint_as_float(raw[8 * g + 4]), __uint_as_float(raw[8 * g + 5]));
      const uint32_t d = pack_bf16x2(__uint_as_float(raw[8 * g + 6]), __uint_as_float(raw[8 * g + 7]));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)(((half * 4 + g) ^ (r & 7)) * 16)), "r"(a),
                   "r"(b), "r"(c), "r"(d)
                   : "memory");
    }
    if (colsum != nullptr) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(raw[k]);
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int k = 0; k < step; ++k) {
          const float keep = up ? v[k + step] : v[k];
          const float send = up ? v[k] : v[k + step];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
      }
      atomicAdd(colsum + half * 32 + lane, v[0]);
    }
  }
}

__global__ void __launch_bounds__(ATB_THREADS, 1)
mhsa_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const MhsaBwdParams p) {
  using L = MhsaBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sPT = smem + L::OFF_PT;
  uint8_t* sDST = smem + L::OFF_DST;
  float* sLse = reinterpret_cast<float*>(smem + L::OFF_LSE);
  float* sDelta = reinterpret_cast<float*>(smem + L::OFF_DELTA);
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* bar_vdo = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;    // S^T, dP^T ready in TMEM          (once per (j,i) pair)
  uint64_t* bar_pd = bar_qk + 3;   // P^T, dS^T written to smem         (128 arrivals per pair)
  uint64_t* bar_acc = bar_qk + 4;  // dV/dK/dQ MMAs of the pair retired (once per pair)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x;
  const int s = blockIdx.y;
  const int row0 = s * ATB_N;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_vdo, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_pd, 256);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tST = tmem_base, tDPT = tmem_base + 128, tDK = tmem_base + 256, tDV = tmem_base + 320,
                 tDQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_qk, 2 * L::TILE);
      for (int b = 0; b < 2; ++b) {
        tma_load_2d(smem + L::OFF_Q + b * 16384, &tmQKV, bar_qk, h * ATB_D, row0 + b * 128);
        tma_load_2d(smem + L::OFF_K + b * 16384, &tmQKV, bar_qk, p.E + h * ATB_D, row0 + b * 128);
      }
      mbar_arrive_expect_tx(bar_vdo, 2 * L::TILE);
      for (int b = 0; b < 2; ++b) {
        tma_load_2d(smem + L::OFF_V + b * 16384, &tmQKV, bar_vdo, 2 * p.E + h * ATB_D, row0 + b * 128);
        tma_load_2d(smem + L::OFF_DO + b * 16384, &tmDO, bar_vdo, h * ATB_D, row0 + b * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aQ = smem_u32(smem + L::OFF_Q), aK = smem_u32(smem + L::OFF_K), aV = smem_u32(smem + L::OFF_V),
                     aDO = smem_u32(smem + L::OFF_DO), aPT = smem_u32(sPT), aDST = smem_u32(sDST);
      const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);   // K-major x K-major
      const uint32_t id_kn = umma_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major
      const uint32_t id_nn = umma_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major
      mbar_wait(bar_qk, 0);
      mbar_wait(bar_vdo, 0);
      tc_fence_after();
      for (int t = 0; t < 4; ++t) {
        const int j = t >> 1, i = t & 1;
        // S^T = K_j Q_i^T ; dP^T = V_j dO_i^T      (K = 64 -> 4 k-steps of 32 B inside the swizzle row)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tST, umma_smem_desc_sw128(aK + j * 16384 + ks * 32, 16, 1024),
                  umma_smem_desc_sw128(aQ + i * 16384 + ks * 32, 16, 1024), id_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tDPT, umma_smem_desc_sw128(aV + j * 16384 + ks * 32, 16, 1024),
                  umma_smem_desc_sw128(aDO + i * 16384 + ks * 32, 16, 1024), id_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_pd, t & 1);
        tc_fence_after();
        // contractions over the 128 queries / keys of the pair: 8 k-steps
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t a_k = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);  // K-major A inside P^T / dS^T
          // dV_j += P^T dO_i
          umma_ss(tDV, umma_smem_desc_sw128(aPT + a_k, 16, 1024),
                  umma_smem_desc_sw128(aDO + i * 16384 + ks * 2048, 8192, 1024), id_kn, (i > 0 || ks > 0) ? 1u : 0u);
          // dK_j += dS^T Q_i
          umma_ss(tDK, umma_smem_desc_sw128(aDST + a_k, 16, 1024),
                  umma_smem_desc_sw128(aQ + i * 16384 + ks * 2048, 8192, 1024), id_kn, (i > 0 || ks > 0) ? 1u : 0u);
          // dQ_i += dS K_j      (A = dS^T tile read MN-major: 64-query chunks 16 KB apart, 16 key-rows = 2048 B)
          umma_ss(tDQ + i * 64, umma_smem_desc_sw128(aDST + ks * 2048, 16384, 1024),
                  umma_smem_desc_sw128(aK + j * 16384 + ks * 2048, 8192, 1024), id_nn, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_acc);
      }
    }
    __syncwarp();
  } else {
    // Two warpgroups share the 128 key rows (TMEM lanes) and split the 128 query columns of each pair in halves;
    // the epilogues are split the same way (dK | dV, dQ_0 | dQ_1).
    const int wg = (warp - 2) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;  // key row within tile j / query row in the epilogues
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    // delta[q] (precomputed, coalesced) and lse2 -> smem   (256 threads <-> 256 query rows)
    {
      const int rr = wg * 128 + r;
      const size_t gi = ((size_t)s * p.H + h) * ATB_N + rr;
      sDelta[rr] = p.delta[gi];
      sLse[rr] = p.lse2[gi];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float c = p.scale_log2;

    for (int t = 0; t < 4; ++t) {
      const int j = t >> 1, i = t & 1;
      mbar_wait(bar_s, t & 1);
      if (t > 0) mbar_wait(bar_acc, (t - 1) & 1);   // previous pair's MMAs no longer read P^T / dS^T smem
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {  // this warpgroup's two 32-query chunks
        const int ch = wg * 2 + cc;
        uint32_t rs[32], rd[32];
        tmem_ld_32x32(tST + lane_sel + ch * 32, rs);
        tmem_ld_32x32(tDPT + lane_sel + ch * 32, rd);
        tmem_wait_ld();
        uint32_t pp[16], dd[16];
        const float* lse = sLse + i * 128 + ch * 32;
        const float* del = sDelta + i * 128 + ch * 32;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx_b(fmaf(__uint_as_float(rs[2 * e]), c, -lse[2 * e]));
          const float p1 = ex2_approx_b(fmaf(__uint_as_float(rs[2 * e + 1]), c, -lse[2 * e + 1]));
          const float d0 = p0 * (__uint_as_float(rd[2 * e]) - del[2 * e]) * p.scale;
          const float d1 = p1 * (__uint_as_float(rd[2 * e + 1]) - del[2 * e + 1]) * p.scale;
          pp[e] = pack_bf16x2(p0, p1);
          dd[e] = pack_bf16x2(d0, d1);
        }
        const int boff = (ch >> 1) * 16384 + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int chunk = (((ch & 1) * 4 + v4) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(sPT + boff + chunk) = make_uint4(pp[4 * v4], pp[4 * v4 + 1], pp[4 * v4 + 2], pp[4 * v4 + 3]);
          *reinterpret_cast<uint4*>(sDST + boff + chunk) = make_uint4(dd[4 * v4], dd[4 * v4 + 1], dd[4 * v4 + 2], dd[4 * v4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pd);
      if (i == 1) {
        // key tile j finished: dK_j (warpgroup 0) and dV_j (warpgroup 1) -> dqkv
        mbar_wait(bar_acc, t & 1);
        tc_fence_after();
        bf16* drow = p.dqkv + ((size_t)row0 + j * 128 + r) * (3 * p.E) + h * ATB_D;
        if (wg == 0) store_tmem_row64(tDK + lane_sel, drow + p.E);
        else         store_tmem_row64(tDV + lane_sel, drow + 2 * p.E);
        tc_fence_before();
      }
    }
    // dQ tiles (bar_acc of the last pair has been waited on above): warpgroup w stores dQ_w
    {
      bf16* drow = p.dqkv + ((size_t)row0 + wg * 128 + r) * (3 * p.E) + h * ATB_D;
      store_tmem_row64(tDQ + wg * 64 + lane_sel, drow);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------------------
// Pipelined, persistent variant (default).  ncu on the kernel above: 53 % of the warp samples wait for an accumulator
// (S^T/dP^T ready, previous pair retired) and 13 % at the CTA's exit barrier, tensor pipe 14 % -- the five contractions
// and the softmax-gradient math of a pair ran strictly one after the other, and every CTA paid its 128 KB load,
// TMEM allocation and epilogue unoverlapped.  Here:
//   * one CTA per SM walks (sequence, head) items; the TMA loads of item n+1 are issued as soon as the last MMA of item n
//     has retired and overlap its dK/dV/dQ epilogue;
//   * the pair loop runs over 64-query sub-tiles s = 0..7 (key tile j = s / 4, queries [64 (s % 4), +64)):
//     S^T_s / dP^T_s (64 fp32 columns each) live in TMEM stage s & 1, so the MMAs of sub-tile s+1 run while warpgroup
//     s & 1 does the softmax-gradient math of sub-tile s (the two warpgroups ping-pong);
//   * P^T goes back to TMEM as the bf16 A operand of dV (over the S^T columns it came from); only dS^T passes through
//     shared memory (K-major A of dK, MN-major A of dQ), four 16 KB sub-blocks that are recycled per key tile;
//   * TMEM: stage 0 [0,128) stage 1 [128,256) | dK [256,320) dV [320,384) dQ_0 [384,448) dQ_1 [448,512).
// ---------------------------------------------------------------------------------------------------------
// Diagnostic timeline (tools/att_trace.py; compiled in only with -DCCD_ATT_TRACE=1, never in the product build): CTA 0 appends
// (role, event, item, sub-tile, clock64) records so that the waiting of every role of the pipelined kernel can be laid out
// on one time axis.  roles: 0 TMA producer, 1 MMA issuer, 2 / 3 softmax-gradient warpgroup 0 / 1 (lane 0 of its first warp).
#ifndef CCD_ATT_TRACE
#define CCD_ATT_TRACE 0
#endif
#if CCD_ATT_TRACE
// Diagnostic build only (tools/att_trace.py).  CTA 0 records (event, item, sub-tile, clock) per role into SHARED memory -- one
// LDS + two STS per record, ~50 cycles; the first version took a global atomicAdd per record (~1000 cycles each, on the single
// issuing thread: the traced kernel ran 2.2x slower than the product and its proportions could not be trusted) -- and copies
// the records to the global buffer when the kernel ends.  4 roles x ATT_TRACE_CAP records x 8 bytes after the product's layout.
constexpr int ATT_TRACE_CAP = 640;
constexpr int ATT_TRACE_BYTES = 4 * ATT_TRACE_CAP * 8 + 16;
__device__ long long* g_att_trace_buf = nullptr;
__device__ unsigned int g_att_trace_n = 0;
__device__ unsigned int g_att_trace_cap = 0;
__device__ __forceinline__ void att_trace_at(uint8_t* base, int role, int ev, int item, int sub) {
  if (blockIdx.x != 0) return;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base) + role;
  const uint32_t i = *cnt;
  if (i < (uint32_t)ATT_TRACE_CAP) {
    uint2* r = reinterpret_cast<uint2*>(base + 16) + role * ATT_TRACE_CAP + i;
    *r = make_uint2((uint32_t)ev | ((uint32_t)sub << 8) | ((uint32_t)item << 16), (uint32_t)clock64());
    *cnt = i + 1;
  }
}
#define ATT_TRACE(role, ev, item, sub) att_trace_at(att_trace_base, role, ev, item, sub)
#define ATT_TRACE_L0(role, ev, item, sub) do { if ((threadIdx.x & 31) == 0) att_trace_at(att_trace_base, role, ev, item, sub); } while (0)
#else
constexpr int ATT_TRACE_BYTES = 0;
#define ATT_TRACE(role, ev, item, sub) ((void)0)
#define ATT_TRACE_L0(role, ev, item, sub) ((void)0)
#endif
// events: 1 wait begin / 2 wait end on bar_done (producer) ; 10/11 bar_qk+bar_vdo, 12/13 bar_pd, 14/15 bar_epi, 16 S/dP issued,
// 17 dV/dK(/dQ) issued (issuer) ; 20/21 bar_free, 22/23 bar_s, 24 math + P/dS written, 25/26 bar_acc, 27 dK/dV/dQ stored (softmax)

struct MhsaBwd2Smem {
  static constexpr int TILE = ATB_N * ATB_D * 2;  // 32 KB
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = TILE;
  static constexpr int OFF_V = 2 * TILE;
  static constexpr int OFF_DO = 3 * TILE;
  static constexpr int OFF_DS = 4 * TILE;             // 4 sub-blocks [128 keys x 64 queries] bf16, 16 KB each
  static constexpr int OFF_LSE = 6 * TILE;            // 2 x 256 f32 (double buffered per item): -lse2
  static constexpr int OFF_DELTA = OFF_LSE + 2048;    // 2 x 256 f32: delta * scale
  static constexpr int OFF_CSUM = OFF_DELTA + 2048;   // [2 (q, v)][8 heads][64] f32 column sums of dQ / dV of this CTA's items
  static constexpr int OFF_BAR = OFF_CSUM + 4096;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

// 1-D bulk copy global -> shared (TMA engine), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sts128_b(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128_b(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(ATB_THREADS, 1)
mhsa_bwd_pipelined_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                          const __grid_constant__ CUtensorMap tmDQKV, const MhsaBwdParams p, int n_items) {
  using L = MhsaBwd2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* bar_qk = bars;           // rows 0..127 of Q, K, V, dO + the lse / delta vectors landed   (1 completion / item)
  uint64_t* bar_vdo = bars + 1;      // rows 128..255 of Q, K, V, dO landed
  uint64_t* bar_done = bars + 2;     // every MMA of the item retired: smem reusable   (1 / item)
  uint64_t* bar_s = bars + 3;        // [2] S^T_s, dP^T_s ready in TMEM stage          (4 / item each)
  uint64_t* bar_pd = bars + 5;       // [2] P^T (TMEM) and dS^T (smem) of sub-tile written, 128 arrivals (4 / item each)
  uint64_t* bar_free = bars + 7;     // [4] dS^T sub-block no longer read by an MMA    (2 / item each)
  uint64_t* bar_acc = bars + 11;     // dK_j, dV_j (and, for j = 1, dQ) complete       (2 / item)
  uint64_t* bar_epi = bars + 12;     // dK_j / dV_j (+dQ) read out of TMEM, 256 arrivals (2 / item)
  uint64_t* bar_stg = bars + 13;     // [2] warpgroup g's staged output tiles have been read by their TMA stores (2 / item each)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_vdo, 1);
    mbar_init(bar_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_pd[i], 128);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bar_free[i], 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_epi, 256);
    mbar_init(&bar_stg[0], 1);
    mbar_init(&bar_stg[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
  }
  float* csum = reinterpret_cast<float*>(smem + L::OFF_CSUM);
#if CCD_ATT_TRACE
  uint8_t* att_trace_base = smem + L::SMEM_BYTES - 1024;       // behind the product layout (the launch adds ATT_TRACE_BYTES)
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(att_trace_base)[threadIdx.x] = 0u;
#endif
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) csum[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_full_base(tmem_slot);
  const uint32_t tDK = tmem_base + 256, tDV = tmem_base + 320, tDQ = tmem_base + 384;
  pdl_wait();                                      // delta / lse vectors and dO of the preceding kernels are complete and visible

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int s = w / p.H, h = w - s * p.H;
        const int row0 = s * ATB_N;
        ATT_TRACE(0, 1, it, 0);
        mbar_wait(bar_done, (it & 1) ^ 1);            // previous item's MMAs no longer read Q/K/V/dO (passes for it = 0)
        ATT_TRACE(0, 2, it, 0);
        // first halves (rows 0..127 of Q, K, V, dO) + the two vectors: everything the first two sub-tiles need
        mbar_arrive_expect_tx(bar_qk, 2 * L::TILE + 2048);
        // -lse2 and -delta*scale of the item's 256 queries (pre-formed by mhsa_delta_kernel), double buffered per item
        bulk_load_1d(smem + L::OFF_LSE + (it & 1) * 1024, p.nlse + (size_t)w * ATB_N, 1024, bar_qk);
        bulk_load_1d(smem + L::OFF_DELTA + (it & 1) * 1024, p.delta + (size_t)w * ATB_N, 1024, bar_qk);
        tma_load_2d(smem + L::OFF_K, &tmQKV, bar_qk, p.E + h * ATB_D, row0);
        tma_load_2d(smem + L::OFF_Q, &tmQKV, bar_qk, h * ATB_D, row0);
        tma_load_2d(smem + L::OFF_V, &tmQKV, bar_qk, 2 * p.E + h * ATB_D, row0);
        tma_load_2d(smem + L::OFF_DO, &tmDO, bar_qk, h * ATB_D, row0);
        // second halves (rows 128..255): first needed by sub-tile 2, i.e. after the first softmax-gradient phase
        mbar_arrive_expect_tx(bar_vdo, 2 * L::TILE);
        tma_load_2d(smem + L::OFF_Q + 16384, &tmQKV, bar_vdo, h * ATB_D, row0 + 128);
        tma_load_2d(smem + L::OFF_DO + 16384, &tmDO, bar_vdo, h * ATB_D, row0 + 128);
        tma_load_2d(smem + L::OFF_K + 16384, &tmQKV, bar_vdo, p.E + h * ATB_D, row0 + 128);
        tma_load_2d(smem + L::OFF_V + 16384, &tmQKV, bar_vdo, 2 * p.E + h * ATB_D, row0 + 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t aQ = smem_u32(smem + L::OFF_Q), aK = smem_u32(smem + L::OFF_K), aV = smem_u32(smem + L::OFF_V),
                     aDO = smem_u32(smem + L::OFF_DO), aDS = smem_u32(smem + L::OFF_DS);
      const uint32_t id_s = umma_idesc_bf16(128, 64, 0, 0);    // S^T / dP^T sub-tile: K-major x K-major, N = 64 queries
      const uint32_t id_kn = umma_idesc_bf16(128, 64, 0, 1);   // dV / dK: A K-major (TMEM or smem), B MN-major
      const uint32_t id_nn = umma_idesc_bf16(128, 64, 1, 1);   // dQ: A MN-major, B MN-major
      // base descriptors, one per (tile, layout); every MMA below advances one of them by a constant (umma_desc_advance): the
      // issuing thread is the critical resource of this kernel -- N = 64 MMAs execute in 32 cycles, so re-deriving two
      // descriptors per MMA (plus the TMEM-address waterfall, see tmem_full_base) made the ISSUE, not the tensor pipe, the limit
      const uint64_t dK_k = umma_smem_desc_sw128(aK, 16, 1024), dQ_k = umma_smem_desc_sw128(aQ, 16, 1024),      // K-major
                     dV_k = umma_smem_desc_sw128(aV, 16, 1024), dDO_k = umma_smem_desc_sw128(aDO, 16, 1024);
      const uint64_t dDO_m = umma_smem_desc_sw128(aDO, 8192, 1024), dQ_m = umma_smem_desc_sw128(aQ, 8192, 1024),  // MN-major
                     dK_m = umma_smem_desc_sw128(aK, 8192, 1024), dDS_m = umma_smem_desc_sw128(aDS, 16384, 1024);
      // sub-tile s: key tile j = s >> 2, queries [64 qs, 64 qs + 64) with qs = s & 3 (row offset qs * 8192 B in Q / dO)
      auto issue_s = [&](int s) {
        const int j = s >> 2, qs = s & 3;
        const uint32_t tS = tmem_base + (s & 1) * 128;
        const uint64_t a0 = umma_desc_advance(dK_k, j * 16384), b0 = umma_desc_advance(dQ_k, qs * 8192);
        const uint64_t a1 = umma_desc_advance(dV_k, j * 16384), b1 = umma_desc_advance(dDO_k, qs * 8192);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tS, umma_desc_advance(a0, ks * 32), umma_desc_advance(b0, ks * 32), id_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tS + 64, umma_desc_advance(a1, ks * 32), umma_desc_advance(b1, ks * 32), id_s, ks > 0 ? 1u : 0u);
        umma_commit(&bar_s[s & 1]);
      };
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        ATT_TRACE(1, 10, it, 0);
        mbar_wait(bar_qk, it & 1);                    // first halves: enough for sub-tiles 0 and 1
        ATT_TRACE(1, 11, it, 0);
        tc_fence_after();
        issue_s(0);
        issue_s(1);
        ATT_TRACE(1, 16, it, 1);
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
          const int g = s & 1, j = s >> 2, qs = s & 3;
          ATT_TRACE(1, 12, it, s);
          mbar_wait(&bar_pd[g], (s >> 1) & 1);
          ATT_TRACE(1, 13, it, s);
          if (qs == 0) {
            ATT_TRACE(1, 14, it, s);
            mbar_wait(bar_epi, j ^ 1);                // dK / dV (dQ) accumulators of the previous key tile / item read out
            ATT_TRACE(1, 15, it, s);
          }
          tc_fence_after();
          const uint32_t tP = tmem_base + g * 128;    // bf16 P^T over the first 32 columns of the stage
          const uint64_t bDO = umma_desc_advance(dDO_m, qs * 8192), bQ = umma_desc_advance(dQ_m, qs * 8192);
          const uint32_t acc0 = qs > 0 ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {            // contraction over the 64 queries of the sub-tile
            // dV_j += P^T dO_s        (dO sub-tile read MN-major: 16 query rows = 2048 B)
            umma_ts(tDV, tP + ks * 8, umma_desc_advance(bDO, ks * 2048), id_kn, ks > 0 ? 1u : acc0);
            // dK_j += dS^T Q_s        (dS^T from TMEM, bf16 pairs over columns [64,96) of the stage)
            umma_ts(tDK, tP + 64 + ks * 8, umma_desc_advance(bQ, ks * 2048), id_kn, ks > 0 ? 1u : acc0);
          }
          if (qs & 1) {
            // dQ_I += dS_I K_j over the 128 keys, I = qs >> 1: A = sub-blocks (qs-1, qs) read MN-major (64-query chunks
            // 16 KB apart, 16 key rows = 2048 B)
            const int I = qs >> 1;
            const uint64_t aDSI = umma_desc_advance(dDS_m, I * 32768), bK = umma_desc_advance(dK_m, j * 16384);
            const uint32_t accq = j > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ss(tDQ + I * 64, umma_desc_advance(aDSI, ks * 2048), umma_desc_advance(bK, ks * 2048), id_nn,
                      ks > 0 ? 1u : accq);
            umma_commit(&bar_free[qs - 1]);
            umma_commit(&bar_free[qs]);
          }
          if (qs == 3) umma_commit(bar_acc);
          if (s == 7) umma_commit(bar_done);
          ATT_TRACE(1, 17, it, s);
          if (s == 0) {                               // sub-tile 2 is the first to read rows 128..255 (Q / dO), then K / V of tile 1
            mbar_wait(bar_vdo, it & 1);
            tc_fence_after();
          }
          if (s + 2 < 8) issue_s(s + 2);              // overwrites stage g: after the dV MMAs above (in-order) and after
        }                                             // warpgroup g finished reading it (bar_pd)
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax-gradient warpgroups: thread = key row of tile j (TMEM lane) =====================
    const int g = (warp - 2) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t sLse = smem_u32(smem + L::OFF_LSE), sDel = smem_u32(smem + L::OFF_DELTA);
    const uint32_t sDS = smem_u32(smem + L::OFF_DS);
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2), sc2 = make_float2(p.scale, p.scale);
    int it = 0;
    int n_epi = 0;                                    // epilogues this warpgroup has issued (two per item)
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int s_idx = w / p.H, h = w - s_idx * p.H;
      const int row0 = s_idx * ATB_N;
      const int par = it & 1;
      mbar_wait(bar_qk, it & 1);                      // the item's -lse2 / -delta*scale vectors have landed (bulk copies)
#pragma unroll 1
      for (int s = g; s < 8; s += 2) {
        const int j = s >> 2, qs = s & 3;
        const uint32_t tS = tmem_base + g * 128 + lane_sel;
        const bool tr = (lane == 0 && q4 == ((2 + 4 * g) & 3));      // lane 0 of the warpgroup's first warp
        if (tr) ATT_TRACE(2 + g, 20, it, s);
        mbar_wait(&bar_free[qs], j ^ 1);              // the dQ / dK MMAs that read this dS^T sub-block have retired
        // ... and so have the TMA stores of the output tiles this warpgroup staged in the dS^T sub-blocks at its last epilogue
        if (qs < 2 && n_epi > 0) mbar_wait(&bar_stg[g], (n_epi - 1) & 1);
        if (tr) { ATT_TRACE(2 + g, 21, it, s); ATT_TRACE(2 + g, 22, it, s); }
        mbar_wait(&bar_s[g], (s >> 1) & 1);
        if (tr) ATT_TRACE(2 + g, 23, it, s);
        tc_fence_after();
        const uint32_t ds_row = sDS + qs * 16384 + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {        // 32 queries at a time
          uint32_t rs[32], rd[32];
          tmem_ld_32x32(tS + half * 32, rs);
          tmem_ld_32x32(tS + 64 + half * 32, rd);
          const uint32_t qoff = (uint32_t)(par * 256 + qs * 64 + half * 32) * 4;
          float4 nl[8], nd[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            nl[e] = lds128_b(sLse + qoff + e * 16);
            nd[e] = lds128_b(sDel + qoff + e * 16);
          }
          tmem_wait_ld();
          uint32_t pp[16], dd[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            // P^T = exp2(S^T c - lse2[q]);  dS^T = P^T (dP^T scale - delta[q] scale)
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(rs[4 * e]), __uint_as_float(rs[4 * e + 1])), c2,
                                         make_float2(nl[e].x, nl[e].y));
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(rs[4 * e + 2]), __uint_as_float(rs[4 * e + 3])), c2,
                                         make_float2(nl[e].z, nl[e].w));
            const float2 p0 = make_float2(ex2_approx_b(x0.x), ex2_approx_b(x0.y));
            const float2 p1 = make_float2(ex2_approx_b(x1.x), ex2_approx_b(x1.y));
            const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(rd[4 * e]), __uint_as_float(rd[4 * e + 1])), sc2,
                                         make_float2(nd[e].x, nd[e].y));
            const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(rd[4 * e + 2]), __uint_as_float(rd[4 * e + 3])), sc2,
                                         make_float2(nd[e].z, nd[e].w));
            const float2 d0 = __fmul2_rn(p0, t0), d1 = __fmul2_rn(p1, t1);
            pp[2 * e] = pack_bf16x2(p0.x, p0.y);
            pp[2 * e + 1] = pack_bf16x2(p1.x, p1.y);
            dd[2 * e] = pack_bf16x2(d0.x, d0.y);
            dd[2 * e + 1] = pack_bf16x2(d1.x, d1.y);
          }
          // P^T -> TMEM (bf16 pairs: 32 queries = 16 columns); the S^T columns it lands on were read in an earlier half
          // or just above (half 0 writes cols [0,16), read range of half 1 is [32,64))
          tmem_st_32x16(tS + half * 16, pp);
          // dS^T -> TMEM as well (over dP^T columns already consumed: half 0 writes cols [64,80), half 1 reads [96,128)): the
          // A operand of dK_j += dS^T Q (TS form) -- one 4 KB shared-memory operand read less per MMA; the shared-memory copy
          // below remains for dQ, which reads the tile MN-major
          tmem_st_32x16(tS + 64 + half * 16, dd);
          // dS^T -> shared memory: row = key r, 16-byte chunk (8 queries) index ^ (r & 7)
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4)
            sts128_b(ds_row + (uint32_t)(((half * 4 + v4) ^ (r & 7)) * 16), dd[4 * v4], dd[4 * v4 + 1], dd[4 * v4 + 2], dd[4 * v4 + 3]);
        }
        tmem_wait_st();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bar_pd[g]);
        if (tr) ATT_TRACE(2 + g, 24, it, s);
        if (qs >= 2) {
          // last sub-tile of this warpgroup in key tile j: once the tile's MMAs retired, write dK_j (warpgroup 0) /
          // dV_j (warpgroup 1); after the second key tile also dQ_0 / dQ_1
          if (tr) ATT_TRACE(2 + g, 25, it, s);
          mbar_wait(bar_acc, j);
          if (tr) ATT_TRACE(2 + g, 26, it, s);
          tc_fence_after();
          // Every MMA of key tile j has retired, so the four dS^T sub-blocks are idle: sub-block 2 + g takes dK_j (g = 0) /
          // dV_j (g = 1), and after the second key tile sub-block g takes dQ rows [128 g, 128 g + 128).  The accumulators are
          // handed back (bar_epi) as soon as they are in shared memory; ONE thread per warpgroup then issues the bulk tensor
          // stores.  Both sub-blocks are next written by THIS warpgroup (qs = g and qs = 2 + g are its sub-tiles): bar_stg[g].
          const bool want_bias = p.dbias != nullptr;
          const uint32_t stage_kv = sDS + (2 + g) * 16384, stage_q = sDS + g * 16384;
          stage_tmem_row64((g == 0 ? tDK : tDV) + lane_sel, stage_kv + r * 128, r, nullptr, lane);
          if (j == 1) stage_tmem_row64(tDQ + g * 64 + lane_sel, stage_q + r * 128, r, want_bias ? csum + h * 64 : nullptr, lane);
          tc_fence_before();
          mbar_arrive(bar_epi);
          fence_proxy_async_smem();
          named_bar_sync(1 + g, 128);
          if (q4 == ((2 + 4 * g) & 3) && lane == 0) {
            tma_store_2d(&tmDQKV, smem + L::OFF_DS + (2 + g) * 16384, (g == 0 ? p.E : 2 * p.E) + h * ATB_D, row0 + j * 128);
            if (j == 1) tma_store_2d(&tmDQKV, smem + L::OFF_DS + g * 16384, h * ATB_D, row0 + g * 128);
            tma_store_commit();
            tma_store_wait_read();
            mbar_arrive(&bar_stg[g]);
          }
          ++n_epi;
          if (tr) ATT_TRACE(2 + g, 27, it, s);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbias != nullptr) {          // flush this CTA's column sums of dQ: the q part [0, E) of the qkv-bias gradient
    for (int i = threadIdx.x; i < p.H * 64; i += blockDim.x) {
      const float v = csum[i];
      if (v != 0.f) atomicAdd(p.dbias + i, v);
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#if CCD_ATT_TRACE
  if (blockIdx.x == 0 && g_att_trace_buf != nullptr) {          // after the __syncthreads above: every role's records are visible
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(att_trace_base);
    for (int role = 0; role < 4; ++role) {
      const uint32_t n = cnt[role];
      uint32_t off = 0;
      for (int r2 = 0; r2 < role; ++r2) off += cnt[r2];
      const uint2* src = reinterpret_cast<const uint2*>(att_trace_base + 16) + role * ATT_TRACE_CAP;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (off + i >= g_att_trace_cap) break;
        long long* dst = g_att_trace_buf + 4 * (size_t)(off + i);
        const uint2 v = src[i];
        dst[0] = ((long long)role << 32) | (v.x & 0xFFu);
        dst[1] = v.x >> 16;
        dst[2] = (v.x >> 8) & 0xFFu;
        dst[3] = v.y;
      }
    }
    if (threadIdx.x == 0) g_att_trace_n = cnt[0] + cnt[1] + cnt[2] + cnt[3];
  }
#endif
}

}  // namespace ccd

using namespace ccd;

// C ABI -- see include/ccd_b200.h
static int g_mhsa_bwd_variant = 1;   // 1 = pipelined persistent kernel (default), 0 = one CTA per (sequence, head)

extern "C" int ccd_colsum_bf16(const void* x, float* out, int rows, int cols, void* stream);
extern "C" int ccd_vecmat_add_f32(const float* v, const float* W, float* out, int rows, int cols, void* stream);

extern "C" int ccd_mhsa_bwd(const void* qkv, const void* o, const void* d_o, const float* lse2, float* delta_ws, void* dqkv,
                            float* dbias_qkv, const float* dproj_bias, const float* w_proj, int S, int H, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!qkv || !o || !d_o || !lse2 || !delta_ws || !dqkv || S <= 0 || H <= 0 || H > 8) return CCD_ERR_ARG;
  if (dbias_qkv != nullptr && g_mhsa_bwd_variant != 1) return CCD_ERR_UNSUPPORTED;   // fused bias gradient: pipelined kernel only
  const int E = H * ATB_D;
  CUtensorMap tmQKV, tmDO, tmDQKV;
  if (!get_tmap_bf16_2d(&tmDQKV, dqkv, (uint64_t)S * ATB_N, (uint64_t)3 * E, (uint64_t)3 * E, 128, 64)) return CCD_ERR_TMAP;
  if (!get_tmap_bf16_2d(&tmQKV, qkv, (uint64_t)S * ATB_N, (uint64_t)3 * E, (uint64_t)3 * E, 128, 64)) return CCD_ERR_TMAP;
  if (!get_tmap_bf16_2d(&tmDO, d_o, (uint64_t)S * ATB_N, (uint64_t)E, (uint64_t)E, 128, 64)) return CCD_ERR_TMAP;
  MhsaBwdParams p;
  p.o = reinterpret_cast<const bf16*>(o);
  p.d_o = reinterpret_cast<const bf16*>(d_o);
  p.lse2 = lse2;
  p.delta = delta_ws;
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.dbias = dbias_qkv;
  p.E = E;
  p.H = H;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(mhsa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        MhsaBwdSmem::SMEM_BYTES));
    attr_set = true;
  }
  const bool pipelined = g_mhsa_bwd_variant == 1;
  float* nlse_ws = pipelined ? delta_ws + (size_t)S * H * ATB_N : nullptr;      // second half of the workspace
  p.nlse = nlse_ws;
  CCD_CUDA_CHECK(launch_pdl(mhsa_delta_kernel, dim3((S * ATB_N + 7) / 8), dim3(256), 0, stream, p.o, p.d_o, delta_ws, lse2, nlse_ws, p.scale,
                            S * ATB_N, E, H));
  if (g_mhsa_bwd_variant == 1) {
    static bool attr2_set = false;
    static int num_sms = 148;
    if (!attr2_set) {
      CCD_CUDA_CHECK(cudaFuncSetAttribute(mhsa_bwd_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          MhsaBwd2Smem::SMEM_BYTES + ATT_TRACE_BYTES));
      int dev = 0;
      CCD_CUDA_CHECK(cudaGetDevice(&dev));
      CCD_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
      attr2_set = true;
    }
    const int n_items = S * H;
    CCD_CUDA_CHECK(launch_pdl(mhsa_bwd_pipelined_kernel, dim3(n_items < num_sms ? n_items : num_sms), dim3(ATB_THREADS),
                              (size_t)(MhsaBwd2Smem::SMEM_BYTES + ATT_TRACE_BYTES), stream, tmQKV, tmDO, tmDQKV, p, n_items));
    if (dbias_qkv != nullptr) {       // v part of the qkv-bias gradient = column sums of d_o (see the comment above the kernel)
      if (dproj_bias != nullptr && w_proj != nullptr) return ccd_vecmat_add_f32(dproj_bias, w_proj, dbias_qkv + 2 * E, E, E, stream_);
      return ccd_colsum_bf16(d_o, dbias_qkv + 2 * E, S * ATB_N, E, stream_);
    }
    return CCD_OK;
  }
  dim3 grid(H, S);
  mhsa_bwd_kernel<<<grid, ATB_THREADS, MhsaBwdSmem::SMEM_BYTES, stream>>>(tmQKV, tmDO, p);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

// debug / A-B switch for the attention backward: 1 = pipelined persistent kernel (default), 0 = first version
extern "C" int ccd_set_mhsa_bwd_variant(int value) {
  g_mhsa_bwd_variant = value ? 1 : 0;
  return CCD_OK;
}

#if CCD_ATT_TRACE
// diagnostic builds only (tools/att_trace.py): buf = device buffer of cap records x 4 int64
extern "C" int ccd_debug_att_trace(long long* buf, unsigned int cap) {
  const unsigned int zero = 0;
  CCD_CUDA_CHECK(cudaMemcpyToSymbol(ccd::g_att_trace_buf, &buf, sizeof(buf)));
  CCD_CUDA_CHECK(cudaMemcpyToSymbol(ccd::g_att_trace_cap, &cap, sizeof(cap)));
  CCD_CUDA_CHECK(cudaMemcpyToSymbol(ccd::g_att_trace_n, &zero, sizeof(zero)));
  return CCD_OK;
}
extern "C" int ccd_debug_att_trace_count(unsigned int* out) {
  CCD_CUDA_CHECK(cudaMemcpyFromSymbol(out, ccd::g_att_trace_n, sizeof(unsigned int)));
  return CCD_OK;
}
#endif
