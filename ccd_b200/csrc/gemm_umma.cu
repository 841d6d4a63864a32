// tcgen05 (UMMA) bf16 GEMM with fused epilogues for the CCD encoder / DINO head linear contractions
// (reference call sites: Dino/modules/vision_transformer.py:59-65 Mlp, :82,:90 Attention.qkv/proj,
//  :126-131 PatchEmbed, :324-328 DINOHead) and their backward (dgrad / wgrad).
//
//   C[M,N] = A[M,K] * B[N,K]^T   bf16 operands, fp32 accumulation in TMEM.
//
// Operand majorness is a runtime flag so the same kernel serves forward (A,B K-major), dgrad (B MN-major: the
// weight [N_out,K_in] read as [K_in-major]) and wgrad (A and B MN-major: activations read transposed) without any
// transposed copies in HBM:
//   K-major  operand X[rows, K]   : memory row-major [rows, K]  (ld = K-extent), TMA box {64 k, rows}
//   MN-major operand X[rows, K]   : memory row-major [K, rows]  (ld = rows-extent), TMA boxes {64 rows, 64 k}
// Tiles land in shared memory through TMA with SWIZZLE_128B and are consumed by tcgen05.mma via smem descriptors.
//
// CTA = 128 x BN output tile, 192 threads: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps 2-5 =
// epilogue (TMEM -> registers -> fused epilogue -> global).  BN=128: 3-stage ring, two CTAs co-resident per SM so
// one CTA's epilogue overlaps the other's main loop.  gridDim.z = split-K slices (fp32 atomic accumulation).
#include "ccd_common.cuh"
#include "tmap.cuh"

namespace ccd {

enum GemmEpi {
  EPI_BF16 = 0,    // out0 bf16 = acc + bias
  EPI_GELU = 1,    // out0 bf16 = acc + bias (pre-activation), out1 bf16 = gelu(out0)
  EPI_RESID = 2,   // out0 f32  = aux_f32[m,n] + acc + bias            (residual stream)
  EPI_F32 = 3,     // out0 f32  = acc + bias   (atomicAdd when split-K)
  EPI_DGELU = 4,   // out0 bf16 = acc * gelu'(aux_bf16[m,n])           (backward through GELU)
  EPI_POS = 5,     // out0 f32  = acc + bias + aux_f32[(m % 256), n]   (patch embed + resampled pos-embed)
  EPI_COUNT = 6
};

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;        // majorness flags
  int kb_per_split;      // k-blocks (of 64) per blockIdx.z
  const float* bias;     // [N] or null
  void* out0;
  void* out1;
  const void* aux;
  int ldc;               // leading dim (elements) of out0/out1/aux
  int atomic;            // EPI_F32: accumulate with atomicAdd
  const float* seq_scale;  // EPI_RESID only: per-sequence (row / 256) scale of the branch (DropPath,
                           // vision_transformer.py:27-36); null = 1
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;  // + barriers + alignment slack
};

template <int EPI>
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int row, int col, const float (&acc)[32]) {
  // 32 consecutive columns of one row; N is a multiple of 8, so validity is decided per group of 8.
  const size_t off = (size_t)row * p.ldc + col;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (col + g * 8 >= p.N) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[g * 8 + j];
    if (p.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + g * 8 + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if constexpr (EPI == EPI_BF16) {
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out0) + off + g * 8) = o;
    } else if constexpr (EPI == EPI_GELU) {
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out0) + off + g * 8) = o;
      // gelu is evaluated on the bf16-rounded pre-activation so that backward (which only has the bf16 copy)
      // differentiates exactly the function that was applied
      float h[8];
      h[0] = bf16lo(o.x); h[1] = bf16hi(o.x); h[2] = bf16lo(o.y); h[3] = bf16hi(o.y);
      h[4] = bf16lo(o.z); h[5] = bf16hi(o.z); h[6] = bf16lo(o.w); h[7] = bf16hi(o.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = gelu_erf(h[j]);
      uint4 o2 = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]),
                            pack_bf16x2(h[6], h[7]));
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out1) + off + g * 8) = o2;
    } else if constexpr (EPI == EPI_RESID) {
      const float* r = reinterpret_cast<const float*>(p.aux) + off + g * 8;
      const float4 r0 = *reinterpret_cast<const float4*>(r);
      const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
      if (p.seq_scale != nullptr) {
        const float sc = __ldg(p.seq_scale + (row >> 8));
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sc;
      }
      float* o = reinterpret_cast<float*>(p.out0) + off + g * 8;
      *reinterpret_cast<float4*>(o) = make_float4(v[0] + r0.x, v[1] + r0.y, v[2] + r0.z, v[3] + r0.w);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4] + r1.x, v[5] + r1.y, v[6] + r1.z, v[7] + r1.w);
    } else if constexpr (EPI == EPI_F32) {
      float* o = reinterpret_cast<float*>(p.out0) + off + g * 8;
      if (p.atomic) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(o + j, v[j]);
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else if constexpr (EPI == EPI_DGELU) {
      const uint4 h = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + off + g * 8);
      v[0] *= dgelu_erf(bf16lo(h.x)); v[1] *= dgelu_erf(bf16hi(h.x));
      v[2] *= dgelu_erf(bf16lo(h.y)); v[3] *= dgelu_erf(bf16hi(h.y));
      v[4] *= dgelu_erf(bf16lo(h.z)); v[5] *= dgelu_erf(bf16hi(h.z));
      v[6] *= dgelu_erf(bf16lo(h.w)); v[7] *= dgelu_erf(bf16hi(h.w));
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out0) + off + g * 8) = o;
    } else if constexpr (EPI == EPI_POS) {
      const float* r = reinterpret_cast<const float*>(p.aux) + (size_t)(row & 255) * p.ldc + col + g * 8;
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(r));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(r + 4));
      float* o = reinterpret_cast<float*>(p.out0) + off + g * 8;
      *reinterpret_cast<float4*>(o) = make_float4(v[0] + r0.x, v[1] + r0.y, v[2] + r0.z, v[3] + r0.w);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4] + r1.x, v[5] + r1.y, v[6] + r1.z, v[7] + r1.w);
    }
  }
}

template <int EPI, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, (BN == 128) ? 2 : 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p_in) {
  using Cfg = GemmCfg<BN>;
  GemmParams p = p_in;
  if (blockIdx.z != 0) p.bias = nullptr;  // split-K: the bias is added by slice 0 only
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GEMM_BM;
  const int n0 = blockIdx.x * BN;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;  // host guarantees nkb >= 1 for every z

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const int k0 = (kb_begin + i) * GEMM_BK;
        if (!p.a_mn) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, p.a_mn, p.b_mn);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // K-major: 16 k-elements = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart.
          // MN-major: 16 k-rows = 2048 B; 64-element MN chunks are separate TMA boxes 8192 B apart.
          const uint64_t da = p.a_mn ? umma_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = p.b_mn ? umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // smem slot free once these MMAs retire
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, raw);
      tmem_wait_ld();
      if (row < p.M && n0 + c < p.N) {
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(raw[j]);
        epilogue_store<EPI>(p, row, n0 + c, acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

template <int EPI, int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int splits,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(gemm_umma_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + GEMM_BM - 1) / GEMM_BM, splits);
  gemm_umma_kernel<EPI, BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

}  // namespace ccd

using namespace ccd;

// C ABI -- see include/ccd_b200.h
extern "C" int ccd_gemm_bf16(const void* A, const void* B, int M, int N, int K, int a_mn, int b_mn, int epi,
                             const float* bias, void* out0, void* out1, const void* aux, const float* seq_scale,
                             int ldc, int splits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || epi < 0 || epi >= EPI_COUNT || !A || !B || !out0) return CCD_ERR_ARG;
  if (ldc <= 0) ldc = N;
  if ((epi == EPI_RESID || epi == EPI_DGELU || epi == EPI_POS) && !aux) return CCD_ERR_ARG;
  if (epi == EPI_GELU && !out1) return CCD_ERR_ARG;
  const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  if (splits > 1 && epi != EPI_F32) return CCD_ERR_ARG;
  int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;  // every z slice gets >= 1 k-block

  constexpr int BN = 128;
  CUtensorMap tmA, tmB;
  bool ok;
  // K-major: memory [rows, K] ; MN-major: memory [K, rows]
  if (!a_mn) ok = get_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, GEMM_BM, 64);
  else       ok = get_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)M, 64, 64);
  if (!ok) return CCD_ERR_TMAP;
  if (!b_mn) ok = get_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN, 64);
  else       ok = get_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)N, 64, 64);
  if (!ok) return CCD_ERR_TMAP;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0; p.kb_per_split = kb_per;
  p.bias = bias; p.out0 = out0; p.out1 = out1; p.aux = aux; p.ldc = ldc; p.atomic = (splits > 1) ? 1 : 0;
  p.seq_scale = seq_scale;
  switch (epi) {
    case EPI_BF16:  return launch_gemm<EPI_BF16, BN>(tmA, tmB, p, splits, stream);
    case EPI_GELU:  return launch_gemm<EPI_GELU, BN>(tmA, tmB, p, splits, stream);
    case EPI_RESID: return launch_gemm<EPI_RESID, BN>(tmA, tmB, p, splits, stream);
    case EPI_F32:   return launch_gemm<EPI_F32, BN>(tmA, tmB, p, splits, stream);
    case EPI_DGELU: return launch_gemm<EPI_DGELU, BN>(tmA, tmB, p, splits, stream);
    case EPI_POS:   return launch_gemm<EPI_POS, BN>(tmA, tmB, p, splits, stream);
  }
  return CCD_ERR_ARG;
}
