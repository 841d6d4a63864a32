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

// Experiment switches (tools/build_variants.py builds extra libraries with -DCCD_DBG_EPI=n; never set in the product build):
//   1 = epilogue computes but never stores, 2 = GELU / GELU' replaced by the identity, 3 = epilogue only drains TMEM
#ifndef CCD_DBG_EPI
#define CCD_DBG_EPI 0
#endif

namespace ccd {

enum GemmEpi {
  EPI_BF16 = 0,    // out0 bf16 = acc + bias
  EPI_GELU = 1,    // out0 bf16 = acc + bias (pre-activation; optional, NULL when no backward follows), out1 bf16 = gelu(out0)
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
  int direct;              // full tiles take the transpose-free epilogue (32-byte aligned rows, no split-K atomics)
};

// Implicit-GEMM ("spatial") operand description for the SegHead convolutions (Dino/modules/segmentor.py:37-95): the
// position dimension of an operand (M for a K-major A tile, K for an MN-major B tile) indexes an N x H x W grid of an
// activation viewed as the 5-D tensor {C, W, P, H, N}; every tap adds an (dy, dx) shift (out-of-bounds = zero padding),
// a parity plane P (transposed convolutions, stride 2) and a channel base.
struct ConvSpec {
  int a_spatial, b_spatial;
  int H, W;                  // grid of the spatial operand's positions
  int n_taps, cpt;           // taps; 64-channel chunks per tap (a_spatial: K = n_taps * cpt * 64)
  int b_tap_cols;            // b_spatial: columns of N per tap (= cpt * 64)
  signed char dy[16], dx[16], par[16];
  short cbase[16];
  int out_rowmap;            // epilogue: 1 = output row (n,a,b) -> (n, 2a+py, 2b+px) of an [N,2H,2W,*] tensor (ConvTranspose s2)
  int py, px;
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

// GELU in the fused epilogues.  nn.GELU() is the exact-erf GELU x Phi(x) (vision_transformer.py:49-65).  The ~25
// instruction erff made the fc1 / fc2-dgrad epilogues instruction-issue bound (profiles/), so Phi is evaluated as
//     Phi(x) ~= sigmoid(x (a + b x^2 + c x^4)),  x clamped to [-8, 8]
// with (a,b,c) a minimax fit against the exact erf form (tools/fit_gelu.py): max |gelu - gelu_exact| = 2.5e-5 and
// max |gelu' - gelu'_exact| = 1.1e-4 over the whole real line -- below the bf16 rounding of the stored activation --
// and the correct e^{-x^2/2}-like relative behaviour in the negative tail.  1 MUFU.EX2 + 1 MUFU.RCP + 8 FMA-pipe ops.
__device__ __forceinline__ float gelu_sigmoid(float x, float& xc_out, float& x2_out) {   // returns sigma(v(x))
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  // -log2(e) * (a + b x^2 + c x^4)
  float pl = fmaf(x2, 1.0142631e-3f, -0.10677572f);      //  -log2e * c ,  -log2e * b
  pl = fmaf(x2, pl, -2.3011214f);                         //  -log2e * a
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xc * pl));
  xc_out = xc;
  x2_out = x2;
  return __fdividef(1.0f, 1.0f + e);
}
__device__ __forceinline__ float gelu_fast(float x) {
#if CCD_DBG_EPI == 2
  return x;
#endif
  float xc, x2;
  return x * gelu_sigmoid(x, xc, x2);
}
__device__ __forceinline__ float dgelu_fast(float x) {      // d/dx [x sigma(v)] = sigma + x sigma (1 - sigma) v'(x)
#if CCD_DBG_EPI == 2
  return x;
#endif
  float xc, x2;
  const float sg = gelu_sigmoid(x, xc, x2);
  float vp = fmaf(x2, -3.5151679e-3f, 0.22203388f);       // 5c , 3b
  vp = fmaf(x2, vp, 1.5950158f);                           // a
  return fmaf(xc * sg * (1.0f - sg), vp, sg);
}

// Packed (fp32x2) forms used by the persistent kernel's epilogues: the same polynomial, evaluated two elements per
// instruction (FMUL2 / FFMA2) with sigma(v) = 1/2 + 1/2 tanh(v/2) on ONE MUFU.TANH instead of MUFU.EX2 + MUFU.RCP, and the
// argument clamp moved onto x^2 (one FMNMX): for |x| > 8 the polynomial is frozen at its value at 8, where sigma has
// long saturated.  7-8.5 issue slots per element instead of 15-16 (the GELU epilogues are instruction-issue / latency
// bound: profiles/).   CCD_GELU_TANH=0 keeps the scalar ex2/rcp form.
#ifndef CCD_GELU_TANH
#define CCD_GELU_TANH 1
#endif
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
#if CCD_DBG_EPI == 2
  return x;
#elif CCD_GELU_TANH
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 pl = __ffma2_rn(x2, f2(-3.5151679e-4f), f2(0.037005647f));     // c/2 , b/2
  pl = __ffma2_rn(x2, pl, f2(0.7975079f));                                // a/2
  const float2 w = __fmul2_rn(x, pl);                                     // v(x) / 2
  const float2 h = __fmul2_rn(x, f2(0.5f));
  const float2 t = make_float2(tanh_approx(w.x), tanh_approx(w.y));
  return __ffma2_rn(h, t, h);                                             // x (1/2 + 1/2 tanh(v/2))
#else
  return make_float2(gelu_fast(x.x), gelu_fast(x.y));
#endif
}
__device__ __forceinline__ float2 dgelu_fast2(float2 x) {   // sigma + x sigma (1 - sigma) v'(x), sigma (1 - sigma) = (1 - t^2) / 4
#if CCD_DBG_EPI == 2
  return x;
#elif CCD_GELU_TANH
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 pl = __ffma2_rn(x2, f2(-3.5151679e-4f), f2(0.037005647f));
  pl = __ffma2_rn(x2, pl, f2(0.7975079f));
  float2 vq = __ffma2_rn(x2, f2(-8.7879198e-4f), f2(0.05550847f));       // v'(x) / 4 :  5c/4 , 3b/4
  vq = __ffma2_rn(x2, vq, f2(0.39875395f));                               // a/4
  const float2 w = __fmul2_rn(x, pl);
  const float2 u = __fmul2_rn(x, vq);
  const float2 t = make_float2(tanh_approx(w.x), tanh_approx(w.y));
  const float2 omt2 = __ffma2_rn(make_float2(-t.x, -t.y), t, f2(1.0f));
  const float2 sg = __ffma2_rn(t, f2(0.5f), f2(0.5f));
  return __ffma2_rn(u, omt2, sg);
#else
  return make_float2(dgelu_fast(x.x), dgelu_fast(x.y));
#endif
}

// Phase 2 of the epilogue: one output row per iteration, lane l owns columns [4l, 4l+4) -> every global access of the
// warp is one contiguous 512 B (fp32) / 256 B (bf16) row segment.
template <int EPI>
__device__ __forceinline__ void epilogue_row(const GemmParams& p, int row, int col, float4 v) {
  const size_t off = (size_t)row * p.ldc + col;
  if constexpr (EPI == EPI_BF16) {
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out0) + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else if constexpr (EPI == EPI_GELU) {
    if (p.out0 != nullptr)       // inference (teacher): only the activation is needed
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out0) + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    const float2 g0 = gelu_fast2(make_float2(v.x, v.y)), g1 = gelu_fast2(make_float2(v.z, v.w));
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out1) + off) = make_uint2(pack_bf16x2(g0.x, g0.y), pack_bf16x2(g1.x, g1.y));
  } else if constexpr (EPI == EPI_RESID) {
    const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + off);
    if (p.seq_scale != nullptr) {
      const float sc = __ldg(p.seq_scale + (row >> 8));
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + off) = make_float4(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
  } else if constexpr (EPI == EPI_F32) {
    float* o = reinterpret_cast<float*>(p.out0) + off;
    if (p.atomic) {
      atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w);
    } else {
      *reinterpret_cast<float4*>(o) = v;
    }
  } else if constexpr (EPI == EPI_DGELU) {
    const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.aux) + off);
    v.x *= dgelu_fast(bf16lo(h.x)); v.y *= dgelu_fast(bf16hi(h.x));
    v.z *= dgelu_fast(bf16lo(h.y)); v.w *= dgelu_fast(bf16hi(h.y));
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out0) + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else if constexpr (EPI == EPI_POS) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + (size_t)(row & 255) * p.ldc + col));
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + off) = make_float4(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
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
    // Phase 1: TMEM -> registers (thread = row) -> this warp's 32 x BN fp32 staging tile in shared memory (the TMA
    //          ring is idle once the accumulator is complete), 16-byte chunks XOR-swizzled by the row so that both
    //          the row-per-thread writes and the row-per-iteration reads hit the minimum number of wavefronts.
    // Phase 2: one row per iteration, lane l <-> columns [4l,4l+4): coalesced global loads/stores + fused epilogue.
    static_assert(BN == 128, "epilogue staging assumes 32 sixteen-byte chunks per row");
    const int q = warp & 3;
    float* stage = reinterpret_cast<float*>(smem + q * (32 * BN * 4));
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int phys = (c * 8 + j) ^ lane;
        *reinterpret_cast<uint4*>(stage + lane * BN + phys * 4) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
      }
    }
    __syncwarp();
    const int col = n0 + 4 * lane;
    if (col < p.N) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const int row_base = m0 + q * 32;
      const int nrows = min(32, p.M - row_base);
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) {
        float4 v = *reinterpret_cast<const float4*>(stage + r * BN + ((lane ^ r) * 4));
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        epilogue_row<EPI>(p, row_base + r, col, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks a list of work items (m-block, n-block, k-slice).  Measured on B200
// (profiles/): with 128x128 tiles and K = 384 the one-tile kernel above spends ~7 us of CTA lifetime on 0.8 us of MMA,
// and a persistent 128x128 version is limited by the SINGLE producer thread (a k-block needs 2-4 cp.async.bulk.tensor
// issues against 256 MMA cycles).  Hence:
//   * 128 x BN tiles with BN in {128, 192, 256} (picked so that BN divides N): 384-512 MMA cycles per k-block,
//   * TWO producer warps (A operand / B operand) streaming k-blocks continuously across tile boundaries,
//   * two TMEM accumulators (2 x BN <= 512 columns) so the MMA issuer runs one tile ahead of the epilogue,
//   * two epilogue warpgroups alternating tiles; per 32-column chunk: TMEM -> regs -> 4 KB swizzled smem transpose ->
//     4 rows x 128 B per warp instruction of global I/O with the fused epilogue (auxiliary operands of the chunk are
//     requested before the transpose).
// 352 threads: warp 0 TMA(A), warp 1 MMA, warp 2 TMA(B), warps 3-6 epilogue group 0, warps 7-10 epilogue group 1.
// ---------------------------------------------------------------------------------------------------------
// Diagnostic timeline (tools/gemm_trace.py; compiled in only with -DCCD_GEMM_TRACE=1, never in the product build): CTA 0 appends
// (role, event, item, aux, clock64) records.  roles: 0 TMA producer A, 1 MMA issuer, 2 TMA producer B, 3 / 7 first warp of
// epilogue warpgroup 0 / 1 (lane 0).  events: 1/2 wait begin / end on an empty ring slot (aux = k-block), 10/11 tmem_empty,
// 12/13 full ring slot (aux = k-block), 14 tile's MMAs issued, 20/21 tmem_full, 22 accumulator handed back, 23 tile stored.
#ifndef CCD_GEMM_TRACE
#define CCD_GEMM_TRACE 0
#endif
#if CCD_GEMM_TRACE
// Low-overhead form: every recording thread (one per role) owns a region of the global buffer and a register counter, so a
// record is one 8-byte fire-and-forget store (the first version took a global atomicAdd per record -- ~1000 cycles each on
// the single producer / issuer threads -- and the traced kernel ran 2x slower than the product).  Region r = role (warp id,
// 0..10), GEMM_TRACE_CAP records of {event | aux << 8 | item << 16, clock32}; counts land in the last 11 words at kernel end.
constexpr unsigned int GEMM_TRACE_CAP = 4096;
__device__ unsigned long long* g_gemm_trace_buf = nullptr;
#define GEMM_TRACE(role, ev, item, aux)                                                                                     \
  do {                                                                                                                      \
    if (gt_buf != nullptr && gt_n < GEMM_TRACE_CAP) {                                                                       \
      gt_buf[(size_t)(role) * GEMM_TRACE_CAP + gt_n] =                                                                      \
          ((unsigned long long)(unsigned int)clock64() << 32) | (unsigned int)((ev) | (((aux) & 0xFF) << 8) | ((item) << 16));     \
      ++gt_n;                                                                                                               \
    }                                                                                                                       \
  } while (0)
#define GEMM_TRACE_DECL                                                                                                     \
  unsigned int gt_n = 0;                                                                                                    \
  unsigned long long* const gt_buf = (blockIdx.x == 0) ? g_gemm_trace_buf : nullptr   /* read once: a register, not a load per record */
#define GEMM_TRACE_FLUSH(role)                                                                                              \
  do {                                                                                                                      \
    if (gt_buf != nullptr) gt_buf[(size_t)11 * GEMM_TRACE_CAP + (role)] = gt_n;                                             \
  } while (0)
#else
#define GEMM_TRACE(role, ev, item, aux) ((void)0)
#define GEMM_TRACE_DECL ((void)0)
#define GEMM_TRACE_FLUSH(role) ((void)0)
#endif

constexpr int PG_THREADS = 352;
constexpr int PG_STAGING = 32 * 32 * 4;                            // 4 KB per epilogue warp: one 32x32 fp32 chunk
template <int BN> struct PgCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;            // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;                 // 16 / 24 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 6 : 4;               // 192 / 160 / 192 KB of operands in flight
  static constexpr int SMEM = STAGES * STAGE_BYTES + 8 * PG_STAGING + 256 + 1024;
};

struct PgWork {
  int n_tiles_n, n_tiles, n_items;   // items = n_tiles * splits, item -> (z = item / n_tiles, tile = item % n_tiles)
};

template <int EPI> struct AuxPack { uint32_t a, b, c, d; };

// explicit shared-space accesses for the epilogue transpose (the generic-pointer form compiles to LD.E / ST.E, which are
// tracked on the long scoreboard like global loads)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

// 256-bit global accesses (LDG.256 / STG.256): one full 32-byte sector per lane
__device__ __forceinline__ void stg256(void* gptr, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ldg256(const void* gptr, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(gptr));
}
__device__ __forceinline__ void ldg256_nc(const void* gptr, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(gptr));
}

// Sum over the 32 lanes of each of a thread's 32 values (thread = row, value j = column j): a transposing butterfly,
// the number of live values halves at every exchange; afterwards lane l holds the total of column l.  31 SHFL.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int EPI>
__device__ __forceinline__ AuxPack<EPI> aux_load(const GemmParams& p, int row, int col) {
  AuxPack<EPI> r{0u, 0u, 0u, 0u};
  if constexpr (EPI == EPI_RESID) {
    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ldc + col);
    r.a = v.x; r.b = v.y; r.c = v.z; r.d = v.w;
  } else if constexpr (EPI == EPI_DGELU) {
    const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.aux) + (size_t)row * p.ldc + col);
    r.a = v.x; r.b = v.y;
  } else if constexpr (EPI == EPI_POS) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.aux) + (size_t)(row & 255) * p.ldc + col));
    r.a = v.x; r.b = v.y; r.c = v.z; r.d = v.w;
  }
  return r;
}

// returns the fp32 values behind what was stored (EPI_DGELU: the column sums of the output are its consumer)
template <int EPI>
__device__ __forceinline__ float4 epilogue_row_aux(const GemmParams& p, int row, int col, float4 v, const AuxPack<EPI>& x) {
#if CCD_DBG_EPI == 1
  if (__float_as_uint(v.x) != 0x7fc12345u) {       // never true for real data: keeps the math, drops the stores
    v.x = v.y + v.z;
    if (__float_as_uint(v.x) != 0x7fc12346u) return v;
  }
#endif
  const size_t off = (size_t)row * p.ldc + col;
  if constexpr (EPI == EPI_RESID) {
    if (p.seq_scale != nullptr) {
      const float sc = __ldg(p.seq_scale + (row >> 8));
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + off) =
        make_float4(v.x + __uint_as_float(x.a), v.y + __uint_as_float(x.b), v.z + __uint_as_float(x.c), v.w + __uint_as_float(x.d));
  } else if constexpr (EPI == EPI_DGELU) {
    const float2 d0 = __fmul2_rn(make_float2(v.x, v.y), dgelu_fast2(make_float2(bf16lo(x.a), bf16hi(x.a))));
    const float2 d1 = __fmul2_rn(make_float2(v.z, v.w), dgelu_fast2(make_float2(bf16lo(x.b), bf16hi(x.b))));
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out0) + off) = make_uint2(pack_bf16x2(d0.x, d0.y), pack_bf16x2(d1.x, d1.y));
    return make_float4(d0.x, d0.y, d1.x, d1.y);
  } else if constexpr (EPI == EPI_POS) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + off) =
        make_float4(v.x + __uint_as_float(x.a), v.y + __uint_as_float(x.b), v.z + __uint_as_float(x.c), v.w + __uint_as_float(x.d));
  } else {
    epilogue_row<EPI>(p, row, col, v);
  }
  return v;
}

// EPI_DGELU with out1 != NULL: out1[col] += sum over rows of the fp32 output (= bias gradient of the layer whose
// pre-activation gradient this GEMM produces; replaces a separate column-sum pass over the [T, 4E] tensor).
// cs = this thread's sums over its 8 rows of the chunk; lanes sharing (lane & 7) hold the same 4 columns.
__device__ __forceinline__ void colsum_flush(float* __restrict__ dst, int col, float4 cs, int sub_row, bool col_ok) {
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
    cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
    cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
  }
  if (sub_row == 0 && col_ok) {
    atomicAdd(dst + col, cs.x);
    atomicAdd(dst + col + 1, cs.y);
    atomicAdd(dst + col + 2, cs.z);
    atomicAdd(dst + col + 3, cs.w);
  }
}

template <int EPI, int BN, bool SPATIAL>
__global__ void __launch_bounds__(PG_THREADS, 1)   // 168 registers: warps are allocated in groups of 4 (12 x 32 x 168 <= 64 K)
gemm_umma_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const GemmParams p_in, const PgWork wk, const ConvSpec cs) {
  using Cfg = PgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 8 * PG_STAGING);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_total = (p_in.K + GEMM_BK - 1) / GEMM_BK;
  GEMM_TRACE_DECL;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);                  // one arrive.expect_tx per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 256);              // both epilogue warpgroups read every accumulator
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  pdl_launch_dependents();
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                      // operands / auxiliary tensors of the previous kernel are complete and visible

  if (warp == 0 || warp == 2) {
    // ===================== TMA producers: warp 0 streams A, warp 2 streams B, continuously over all items =====================
    if (lane == 0) {
      const bool is_a = (warp == 0);
      uint32_t kbc = 0;   // running k-block counter -> ring slot / phase
      for (int item = blockIdx.x; item < wk.n_items; item += gridDim.x) {
        const int z = item / wk.n_tiles, tile = item - z * wk.n_tiles;
        const int m0 = (tile / wk.n_tiles_n) * GEMM_BM, n0 = (tile % wk.n_tiles_n) * BN;
        const int kb_begin = z * p_in.kb_per_split;
        const int nkb = min(kb_total, kb_begin + p_in.kb_per_split) - kb_begin;
        for (int i = 0; i < nkb; ++i, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          GEMM_TRACE(warp, 1, item, i);
          mbar_wait(&empty_bar[s], ph ^ 1);
          GEMM_TRACE(warp, 2, item, i);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = (kb_begin + i) * GEMM_BK;
          if (is_a) {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES);
            if (SPATIAL && cs.a_spatial) {
              // K-major A tile = 128 grid positions starting at m0, shifted by the tap of this k-block
              const int kb = kb_begin + i;
              const int tap = kb / cs.cpt, cc = kb - tap * cs.cpt;
              const int hw = cs.H * cs.W;
              const int n_img = m0 / hw, rem = m0 - n_img * hw;
              const int y0 = rem / cs.W, x0 = rem - y0 * cs.W;
              tma_load_5d(sa, &tmA, &full_bar[s], cs.cbase[tap] + cc * 64, x0 + cs.dx[tap], cs.par[tap], y0 + cs.dy[tap], n_img);
            } else if (!p_in.a_mn) {
              tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
            } else {
              tma_load_2d(sa, &tmA, &full_bar[s], m0, k0);
              tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES);
            if (SPATIAL && cs.b_spatial) {
              // MN-major B: 64 grid positions (k-block) x BN columns = (tap, channels) ; one box per 64 channels
              const int tap = n0 / cs.b_tap_cols, c0 = n0 - tap * cs.b_tap_cols;
              const int hw = cs.H * cs.W;
              const int n_img = k0 / hw, rem = k0 - n_img * hw;
              const int y0 = rem / cs.W, x0 = rem - y0 * cs.W;
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_5d(sb + j * 8192, &tmB, &full_bar[s], cs.cbase[tap] + c0 + 64 * j, x0 + cs.dx[tap], cs.par[tap],
                            y0 + cs.dy[tap], n_img);
            } else if (!p_in.b_mn) {
              tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
            }
          }
        }
      }
      GEMM_TRACE_FLUSH(warp);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, p_in.a_mn, p_in.b_mn);
      uint32_t kbc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < wk.n_items; item += gridDim.x, ++it) {
        const int z = item / wk.n_tiles;
        const int kb_begin = z * p_in.kb_per_split;
        const int nkb = min(kb_total, kb_begin + p_in.kb_per_split) - kb_begin;
        const int acc = it & 1;
        GEMM_TRACE(1, 10, item, 0);
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);         // epilogue drained this accumulator
        GEMM_TRACE(1, 11, item, 0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          GEMM_TRACE(1, 12, item, i);
          mbar_wait(&full_bar[s], ph);
          GEMM_TRACE(1, 13, item, i);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = p_in.a_mn ? umma_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = p_in.b_mn ? umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_ss(d_tmem, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
        GEMM_TRACE(1, 14, item, 0);
      }
      GEMM_TRACE_FLUSH(1);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: BOTH warpgroups work on every tile (group g takes column chunks [g, g+1) * BN/64) =====================
    // With one warpgroup per tile the accumulator was held for the whole epilogue E of a tile and the tile period was
    // (M + E) / 2 with E ~ 3 M (ncu: the epilogue warps spent 20 % of their time waiting for the next accumulator while
    // the MMA issuer waited for them).  Two warps per scheduler on the same tile halve the time an accumulator is held
    // and double the latency hiding: period = max(M, E / 2).
    const int g = (warp - 3) >> 2;
    constexpr int CHUNKS_PER_GROUP = BN / 64;
    const int q = warp & 3;                                       // TMEM lane quarter this warp may read
    float* stage = reinterpret_cast<float*>(staging + (warp - 3) * PG_STAGING);
    int it = 0;
    for (int item = blockIdx.x; item < wk.n_items; item += gridDim.x, ++it) {
      GemmParams p = p_in;
      const int z = item / wk.n_tiles, tile = item - z * wk.n_tiles;
      if (z != 0) p.bias = nullptr;
      const int m0 = (tile / wk.n_tiles_n) * GEMM_BM, n0 = (tile % wk.n_tiles_n) * BN;
      const int acc = it & 1;
      const int row_base = m0 + q * 32;
      const int nrows = max(0, min(32, p.M - row_base));
      const int sub_row = lane >> 3, sub_chunk = lane & 7;          // phase 2: 4 rows x 8 sixteen-byte chunks per instruction
      // Pull this tile's auxiliary operand (residual stream / saved pre-activation) into L2 while the MMAs of the tile
      // are still running: the per-chunk loads below then pay L2 instead of HBM latency.  Thread <-> row.
      if constexpr (EPI == EPI_RESID || EPI == EPI_DGELU) {
        constexpr int ES = (EPI == EPI_RESID) ? 4 : 2;
        const int prow = row_base + lane;
        if (prow < p.M) {
          const int nb = n0 + g * (BN / 2);
          const char* base = reinterpret_cast<const char*>(p.aux) + ((size_t)prow * p.ldc + nb) * ES;
          const int bytes = max(0, min(BN / 2, p.N - nb)) * ES;
          for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
        }
      }
      if (p.direct && (row_base + 32 <= p.M) && (n0 + BN <= p.N)) {
        // ---- transpose-free epilogue for full tiles: thread = row, straight from the tcgen05.ld registers ----
        // Every lane owns 32 consecutive columns of its own output row per chunk: 64 B (bf16) / 128 B (fp32), moved with
        // 256-bit global accesses (one full 32-byte sector per lane: tools/ubench/st_rows.cu measures 5.3 TB/s for this
        // pattern against 5.8 TB/s for row-contiguous warps).  The staging-buffer transpose it replaces cost 16 shared-memory
        // wavefront-instructions per chunk and warp, i.e. ~2000 cycles of the SM's shared-memory pipe per 128x256 tile --
        // the same pipe tcgen05.mma reads its operands through (2300 cycles per tile at K = 384): with the transpose the
        // tile period was 7400 cycles against 4100 for the main loop alone (tools/ubench/mma_rate.cu).
        const int row = row_base + lane;
        int orow = row;
        if (SPATIAL && cs.out_rowmap) {          // ConvTranspose2d stride 2: scatter to the (py, px) parity positions
          const int hw = cs.H * cs.W;
          const int n_img = orow / hw, rem = orow - n_img * hw;
          const int a = rem / cs.W, b = rem - a * cs.W;
          orow = (n_img * 2 * cs.H + 2 * a + cs.py) * (2 * cs.W) + 2 * b + cs.px;
        }
        const int nb = n0 + g * (BN / 2);        // first column of this group's half of the tile
        const size_t ooff = (size_t)orow * p.ldc + nb;
        const size_t aoff = (EPI == EPI_POS) ? (size_t)(row & 255) * p.ldc + nb : (size_t)row * p.ldc + nb;
        // the group's BN/2 bias values -> this warp's staging buffer (read back as warp-wide broadcasts)
        const uint32_t bias_s = smem_u32(stage);
        if (4 * lane < BN / 2) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + nb + 4 * lane));
          sts128(bias_s + (uint32_t)lane * 16u, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
        }
        float sc = 1.0f;
        if constexpr (EPI == EPI_RESID) {
          if (p.seq_scale != nullptr) sc = __ldg(p.seq_scale + (row >> 8));
        }
        __syncwarp();
        if (lane == 0 && q == 3) GEMM_TRACE(warp, 20, item, 1);
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        if (lane == 0 && q == 3) GEMM_TRACE(warp, 21, item, 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < CHUNKS_PER_GROUP; ++cc) {
          constexpr int AUXW = (EPI == EPI_RESID || EPI == EPI_POS) ? 32 : (EPI == EPI_DGELU) ? 16 : 1;
          uint32_t ax[AUXW];
          if constexpr (EPI == EPI_RESID) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ldg256(reinterpret_cast<const float*>(p.aux) + aoff + cc * 32 + k * 8, ax + 8 * k);
          } else if constexpr (EPI == EPI_POS) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ldg256_nc(reinterpret_cast<const float*>(p.aux) + aoff + cc * 32 + k * 8, ax + 8 * k);
          } else if constexpr (EPI == EPI_DGELU) {
#pragma unroll
            for (int k = 0; k < 2; ++k) ldg256(reinterpret_cast<const bf16*>(p.aux) + aoff + cc * 32 + k * 16, ax + 8 * k);
          }
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + (g * CHUNKS_PER_GROUP + cc) * 32), raw);
          tmem_wait_ld();
          if (cc == CHUNKS_PER_GROUP - 1) {      // this group's half of the accumulator is read: hand it back to the MMA issuer
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (lane == 0 && q == 3) GEMM_TRACE(warp, 22, item, 1);
          }
#if CCD_DBG_EPI == 3
          if (raw[0] != 0x7fc12345u) continue;
#endif
          float2 f[16];                          // acc + bias, column pairs
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = lds128(bias_s + (uint32_t)(cc * 128 + j * 16));
            f[2 * j] = __fadd2_rn(make_float2(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1])), make_float2(b.x, b.y));
            f[2 * j + 1] = __fadd2_rn(make_float2(__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])), make_float2(b.z, b.w));
          }
          if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU) {
            uint32_t o[16];
            if (EPI == EPI_BF16 || p.out0 != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(f[j].x, f[j].y);
              bf16* dst = reinterpret_cast<bf16*>(p.out0) + ooff + cc * 32;
              stg256(dst, o);
              stg256(dst + 16, o + 8);
            }
            if constexpr (EPI == EPI_GELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 gl = gelu_fast2(f[j]);
                o[j] = pack_bf16x2(gl.x, gl.y);
              }
              bf16* dst = reinterpret_cast<bf16*>(p.out1) + ooff + cc * 32;
              stg256(dst, o);
              stg256(dst + 16, o + 8);
            }
          } else if constexpr (EPI == EPI_DGELU) {
            uint32_t o[16];
            float d[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 r = __fmul2_rn(f[j], dgelu_fast2(make_float2(bf16lo(ax[j]), bf16hi(ax[j]))));
              o[j] = pack_bf16x2(r.x, r.y);
              d[2 * j] = r.x;
              d[2 * j + 1] = r.y;
            }
            bf16* dst = reinterpret_cast<bf16*>(p.out0) + ooff + cc * 32;
            stg256(dst, o);
            stg256(dst + 16, o + 8);
            if (p.out1 != nullptr) {             // bias gradient of the layer below: column sums of the fp32 result
              const float tot = warp_colsum32(d, lane);
              atomicAdd(reinterpret_cast<float*>(p.out1) + nb + cc * 32 + lane, tot);
            }
          } else {                               // fp32 outputs: EPI_RESID / EPI_POS / EPI_F32 (no split-K)
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float2 r = f[j];
              if constexpr (EPI == EPI_RESID) {
                r = __ffma2_rn(r, make_float2(sc, sc), make_float2(__uint_as_float(ax[2 * j]), __uint_as_float(ax[2 * j + 1])));
              } else if constexpr (EPI == EPI_POS) {
                r = __fadd2_rn(r, make_float2(__uint_as_float(ax[2 * j]), __uint_as_float(ax[2 * j + 1])));
              }
              o[2 * j] = __float_as_uint(r.x);
              o[2 * j + 1] = __float_as_uint(r.y);
            }
            float* dst = reinterpret_cast<float*>(p.out0) + ooff + cc * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) stg256(dst + 8 * k, o + 8 * k);
          }
        }
        __syncwarp();                            // the bias staging is rewritten for the next tile
        if (lane == 0 && q == 3) GEMM_TRACE(warp, 23, item, 1);
        continue;
      }
      if (lane == 0 && q == 3) GEMM_TRACE(warp, 20, item, 0);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      if (lane == 0 && q == 3) GEMM_TRACE(warp, 21, item, 0);
      tc_fence_after();
      // Full tiles (every row and column valid) take a branch-free path: all eight staging loads of a chunk are issued
      // back to back, then the 32 independent epilogue chains, then the stores.  (The masked path below serialises
      // load -> math -> store per row group; measured with ncu source counters it made the GELU / GELU' / residual
      // epilogues latency bound at ~1600 cycles per 32-column chunk, 4x the MMA time of the tile.)
      const bool full = (row_base + 32 <= p.M) && (n0 + BN <= p.N);
      const uint32_t stage_s = smem_u32(stage);
      const uint32_t st_wr = stage_s + (uint32_t)lane * 128u;                      // thread = row `lane` of the chunk
      const uint32_t st_rd = stage_s + (uint32_t)sub_row * 128u + (uint32_t)((sub_chunk ^ (sub_row & 7)) * 16);
      const int c_begin = g * CHUNKS_PER_GROUP, c_end = c_begin + CHUNKS_PER_GROUP;
      // bias of the next chunk is requested one chunk ahead (its global-load latency was exposed in every chunk)
      float4 b4_next = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full && p.bias != nullptr) b4_next = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c_begin * 32 + 4 * sub_chunk));
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col = n0 + c * 32 + 4 * sub_chunk;
        if (full) {
          AuxPack<EPI> ax[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ax[i] = aux_load<EPI>(p, row_base + i * 4 + sub_row, col);
          const float4 b4 = b4_next;
          if (p.bias != nullptr && c + 1 < c_end) b4_next = __ldg(reinterpret_cast<const float4*>(p.bias + col + 32));
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), raw);
          tmem_wait_ld();
          if (c == c_end - 1) {                  // this group's half of the accumulator is read: hand it back to the MMA issuer
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (lane == 0 && q == 3) GEMM_TRACE(warp, 22, item, 0);
          }
#if CCD_DBG_EPI == 3
          if (raw[0] != 0x7fc12345u) continue;
#endif
#pragma unroll
          for (int j = 0; j < 8; ++j)            // 16-byte chunk j of row `lane` lands at j ^ (lane & 7)
            sts128(st_wr + (uint32_t)((j ^ (lane & 7)) * 16), raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
          __syncwarp();
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {          // rows r = 4 i + sub_row: (r & 7) = ((4 i) & 7) | sub_row -> chunk index ^ 4 on odd i
            v[i] = lds128((st_rd + (uint32_t)i * 512u) ^ ((i & 1) ? 64u : 0u));
            v[i].x += b4.x; v[i].y += b4.y; v[i].z += b4.z; v[i].w += b4.w;
          }
          __syncwarp();                          // the 4 KB transpose buffer is rewritten by the next chunk
          float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            int orow = row_base + i * 4 + sub_row;
            if (SPATIAL && cs.out_rowmap) {      // ConvTranspose2d stride 2: scatter to the (py, px) parity positions
              const int hw = cs.H * cs.W;
              const int n_img = orow / hw, rem = orow - n_img * hw;
              const int a = rem / cs.W, b = rem - a * cs.W;
              orow = (n_img * 2 * cs.H + 2 * a + cs.py) * (2 * cs.W) + 2 * b + cs.px;
            }
            const float4 o4 = epilogue_row_aux<EPI>(p, orow, col, v[i], ax[i]);
            if constexpr (EPI == EPI_DGELU) { csum.x += o4.x; csum.y += o4.y; csum.z += o4.z; csum.w += o4.w; }
          }
          if constexpr (EPI == EPI_DGELU) {
            if (p.out1 != nullptr) colsum_flush(reinterpret_cast<float*>(p.out1), col, csum, sub_row, true);
          }
          continue;
        }
        const bool col_ok = col < p.N;
        // auxiliary operands + bias of this chunk are requested first: their latency hides behind the transpose
        AuxPack<EPI> ax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col_ok && i * 4 + sub_row < nrows) ax[i] = aux_load<EPI>(p, row_base + i * 4 + sub_row, col);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), raw);
        tmem_wait_ld();
        if (c == c_end - 1) {                    // this group's half of the accumulator is read: hand it back to the MMA issuer
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)              // thread = row `lane`; 16-byte chunk j lands at j ^ (row & 7)
          *reinterpret_cast<uint4*>(stage + lane * 32 + ((j ^ (lane & 7)) * 4)) = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
        __syncwarp();
        float4 csum_m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + sub_row;
          if (col_ok && r < nrows) {
            float4 v = *reinterpret_cast<const float4*>(stage + r * 32 + ((sub_chunk ^ (r & 7)) * 4));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            int orow = row_base + r;
            if (SPATIAL && cs.out_rowmap) {      // ConvTranspose2d stride 2: scatter to the (py, px) parity positions
              const int hw = cs.H * cs.W;
              const int n_img = orow / hw, rem = orow - n_img * hw;
              const int a = rem / cs.W, b = rem - a * cs.W;
              orow = (n_img * 2 * cs.H + 2 * a + cs.py) * (2 * cs.W) + 2 * b + cs.px;
            }
            const float4 o4 = epilogue_row_aux<EPI>(p, orow, col, v, ax[i]);
            if constexpr (EPI == EPI_DGELU) { csum_m.x += o4.x; csum_m.y += o4.y; csum_m.z += o4.z; csum_m.w += o4.w; }
          }
        }
        if constexpr (EPI == EPI_DGELU) {
          if (p.out1 != nullptr) colsum_flush(reinterpret_cast<float*>(p.out1), col, csum_m, sub_row, col_ok);
        }
        __syncwarp();                            // the 4 KB transpose buffer is rewritten by the next chunk
      }
      if (lane == 0 && q == 3) GEMM_TRACE(warp, 23, item, 0);
    }
    if (lane == 0 && q == 3) GEMM_TRACE_FLUSH(warp);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int g_gemm_variant = 1;   // 1 = persistent (default), 0 = one tile per CTA
static int g_gemm_epilogue = 1;  // full-tile epilogue: 1 = per-shape choice (default), 0 = shared-memory transpose everywhere, 2 = transpose-free everywhere

template <int EPI, int BN, bool SPATIAL>
static int launch_gemm_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int splits,
                                  cudaStream_t stream, const ConvSpec& cs) {
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(gemm_umma_persistent_kernel<EPI, BN, SPATIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PgCfg<BN>::SMEM));
    int dev = 0;
    CCD_CUDA_CHECK(cudaGetDevice(&dev));
    CCD_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  PgWork wk;
  wk.n_tiles_n = (p.N + BN - 1) / BN;
  wk.n_tiles = wk.n_tiles_n * ((p.M + GEMM_BM - 1) / GEMM_BM);
  wk.n_items = wk.n_tiles * splits;
  const int grid = wk.n_items < num_sms ? wk.n_items : num_sms;
  CCD_CUDA_CHECK(launch_pdl(gemm_umma_persistent_kernel<EPI, BN, SPATIAL>, dim3(grid), dim3(PG_THREADS), (size_t)PgCfg<BN>::SMEM, stream, tmA,
                            tmB, p, wk, cs));
  return CCD_OK;
}

template <int EPI>
static int dispatch_gemm_persistent(const void* A, const void* B, const GemmParams& p, int splits, cudaStream_t stream) {
  // BN is the widest of {256, 192, 128} that divides N (ragged N falls back to 128 with column masking)
  const int bn = (p.N % 256 == 0) ? 256 : (p.N % 192 == 0) ? 192 : 128;
  CUtensorMap tmA, tmB;
  bool ok;
  if (!p.a_mn) ok = get_tmap_bf16_2d(&tmA, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K, GEMM_BM, 64);
  else         ok = get_tmap_bf16_2d(&tmA, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.M, 64, 64);
  if (!ok) return CCD_ERR_TMAP;
  if (!p.b_mn) ok = get_tmap_bf16_2d(&tmB, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, (uint32_t)bn, 64);
  else         ok = get_tmap_bf16_2d(&tmB, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.N, 64, 64);
  if (!ok) return CCD_ERR_TMAP;
  ConvSpec cs{};
  if (bn == 256) return launch_gemm_persistent<EPI, 256, false>(tmA, tmB, p, splits, stream, cs);
  if (bn == 192) return launch_gemm_persistent<EPI, 192, false>(tmA, tmB, p, splits, stream, cs);
  return launch_gemm_persistent<EPI, 128, false>(tmA, tmB, p, splits, stream, cs);
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
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || epi < 0 || epi >= EPI_COUNT || !A || !B || (!out0 && epi != EPI_GELU)) return CCD_ERR_ARG;
  if (ldc <= 0) ldc = N;
  if ((epi == EPI_RESID || epi == EPI_DGELU || epi == EPI_POS) && !aux) return CCD_ERR_ARG;
  if (epi == EPI_GELU && !out1) return CCD_ERR_ARG;
  if (epi == EPI_DGELU && out1 && g_gemm_variant != 1) return CCD_ERR_UNSUPPORTED;   // fused column sums: persistent kernel only
  const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  if (splits > 1 && epi != EPI_F32) return CCD_ERR_ARG;
  int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;  // every z slice gets >= 1 k-block

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0; p.kb_per_split = kb_per;
  p.bias = bias; p.out0 = out0; p.out1 = out1; p.aux = aux; p.ldc = ldc; p.atomic = (splits > 1) ? 1 : 0;
  p.seq_scale = seq_scale;
  {
    // transpose-free epilogue: 256-bit accesses need 32-byte aligned row segments of every output / auxiliary tensor
    const size_t es0 = (epi == EPI_RESID || epi == EPI_F32 || epi == EPI_POS) ? 4 : 2;
    const size_t esa = (epi == EPI_DGELU) ? 2 : 4;
    // Measured per shape (tools/kbench.py, ViT-Small batch 256): the thread-per-row epilogue wins where one bf16 output
    // (or a long-K fp32 residual tile) leaves the LSU tag stage idle -- 170 vs 178 us for fc1+GELU without the saved
    // pre-activation, 1-2 % on the plain bf16 GEMMs and the K = 1536 residual GEMM -- and loses where two tensors per tile
    // go through sector-per-lane accesses (GELU with saved pre-activation, GELU', K = 384 residual GEMM: 10-28 % slower).
    bool ok = !p.atomic && (g_gemm_epilogue == 2 ||
                            (g_gemm_epilogue == 1 && (epi == EPI_BF16 || (epi == EPI_GELU && out0 == nullptr) || (epi == EPI_RESID && K >= 1024))));
    ok = ok && (((size_t)ldc * es0) % 32 == 0) && (out0 == nullptr || ((uintptr_t)out0 % 32) == 0);
    if (epi == EPI_GELU) ok = ok && ((uintptr_t)out1 % 32) == 0;
    if (epi == EPI_RESID || epi == EPI_DGELU || epi == EPI_POS) ok = ok && (((size_t)ldc * esa) % 32 == 0) && ((uintptr_t)aux % 32) == 0;
    p.direct = ok ? 1 : 0;
  }
  if (g_gemm_variant == 1) {
    switch (epi) {
      case EPI_BF16:  return dispatch_gemm_persistent<EPI_BF16>(A, B, p, splits, stream);
      case EPI_GELU:  return dispatch_gemm_persistent<EPI_GELU>(A, B, p, splits, stream);
      case EPI_RESID: return dispatch_gemm_persistent<EPI_RESID>(A, B, p, splits, stream);
      case EPI_F32:   return dispatch_gemm_persistent<EPI_F32>(A, B, p, splits, stream);
      case EPI_DGELU: return dispatch_gemm_persistent<EPI_DGELU>(A, B, p, splits, stream);
      case EPI_POS:   return dispatch_gemm_persistent<EPI_POS>(A, B, p, splits, stream);
    }
  }
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

// Implicit-GEMM convolution entry (SegHead, Dino/modules/segmentor.py:37-95) on the persistent tcgen05 kernel.
//   spatial_operand = 1: A is the activation (K-major, 128-position tiles): C[M=positions, N] = sum_taps A_shift(tap) * B
//                        (conv3x3 / ConvTranspose forward and data gradients); B is a plain K-major [N, K] weight matrix.
//   spatial_operand = 2: B is the activation (MN-major, K = positions): C[M, N=(tap, channel)] = A^T-style wgrad;
//                        A is a plain MN-major [K=positions, M] matrix (the output gradient).
// sp = the spatial activation, viewed as {C_total, W, P, H, n_img} (P = 2: row/column parity planes of an
// [n_img, 2H, 2W, C_total/2] tensor).  taps_host = int[n_taps][4] = (dy, dx, parity_plane, channel_base).
extern "C" int ccd_conv_gemm(const void* sp, const void* other, int M, int N, int K, int epi, const float* bias, void* out0,
                             int ldc, int splits, int spatial_operand, int H, int W, int C_total, int P, int n_img, int n_taps,
                             const int* taps_host, int cols_per_tap, int rowmap, int py, int px, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!sp || !other || !out0 || !taps_host || M <= 0 || N <= 0 || K <= 0 || (N & 7) || n_taps < 1 || n_taps > 16) return CCD_ERR_ARG;
  if (epi != EPI_BF16 && epi != EPI_F32) return CCD_ERR_ARG;
  if (spatial_operand != 1 && spatial_operand != 2) return CCD_ERR_ARG;
  if (W > 128 && (W % 128)) return CCD_ERR_ARG;
  if ((H * W) % 128 || (cols_per_tap % 64) || (C_total & 7)) return CCD_ERR_ARG;
  if (ldc <= 0) ldc = N;
  const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  if (splits > 1 && epi != EPI_F32) return CCD_ERR_ARG;
  int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;

  ConvSpec cs{};
  cs.a_spatial = spatial_operand == 1;
  cs.b_spatial = spatial_operand == 2;
  cs.H = H; cs.W = W; cs.n_taps = n_taps;
  cs.cpt = cols_per_tap / 64;
  cs.b_tap_cols = cols_per_tap;
  for (int t = 0; t < n_taps; ++t) {
    cs.dy[t] = (signed char)taps_host[4 * t + 0];
    cs.dx[t] = (signed char)taps_host[4 * t + 1];
    cs.par[t] = (signed char)taps_host[4 * t + 2];
    cs.cbase[t] = (short)taps_host[4 * t + 3];
  }
  cs.out_rowmap = rowmap; cs.py = py; cs.px = px;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.kb_per_split = kb_per;
  p.bias = bias; p.out0 = out0; p.out1 = nullptr; p.aux = nullptr; p.ldc = ldc; p.atomic = (splits > 1) ? 1 : 0;
  p.seq_scale = nullptr;
  p.direct = (!p.atomic && (g_gemm_epilogue == 2 || (g_gemm_epilogue == 1 && epi == EPI_BF16)) &&
              (((size_t)ldc * (epi == EPI_F32 ? 4 : 2)) % 32 == 0) && ((uintptr_t)out0 % 32) == 0) ? 1 : 0;

  int bn;
  CUtensorMap tmA, tmB;
  const uint64_t dims[5] = {(uint64_t)C_total, (uint64_t)W, (uint64_t)P, (uint64_t)H, (uint64_t)n_img};
  const uint64_t strides[4] = {(uint64_t)C_total, (uint64_t)W * C_total, (uint64_t)P * W * C_total, (uint64_t)H * P * W * C_total};
  if (cs.a_spatial) {
    if (K != n_taps * cols_per_tap || M != n_img * H * W) return CCD_ERR_ARG;
    p.a_mn = 0; p.b_mn = 0;
    bn = (N % 256 == 0) ? 256 : (N % 192 == 0) ? 192 : 128;
    const uint32_t bw = W >= 128 ? 128 : (uint32_t)W;
    if (!make_tmap_bf16_5d(&tmA, sp, dims, strides, bw, 128 / bw)) return CCD_ERR_TMAP;
    if (!get_tmap_bf16_2d(&tmB, other, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)bn, 64)) return CCD_ERR_TMAP;
  } else {
    if (N != n_taps * cols_per_tap || K != n_img * H * W) return CCD_ERR_ARG;
    p.a_mn = 1; p.b_mn = 1;
    // the N tile must not straddle two taps
    bn = (cols_per_tap % 256 == 0) ? 256 : (cols_per_tap % 192 == 0) ? 192 : (cols_per_tap % 128 == 0) ? 128 : 0;
    if (bn == 0) return CCD_ERR_ARG;
    const uint32_t bw = W >= 64 ? 64 : (uint32_t)W;
    if (!get_tmap_bf16_2d(&tmA, other, (uint64_t)K, (uint64_t)M, (uint64_t)M, 64, 64)) return CCD_ERR_TMAP;
    if (!make_tmap_bf16_5d(&tmB, sp, dims, strides, bw, 64 / bw)) return CCD_ERR_TMAP;
  }
#define CCD_CONV_LAUNCH(E)                                                                    \
  do {                                                                                        \
    if (bn == 256) return launch_gemm_persistent<E, 256, true>(tmA, tmB, p, splits, stream, cs); \
    if (bn == 192) return launch_gemm_persistent<E, 192, true>(tmA, tmB, p, splits, stream, cs); \
    return launch_gemm_persistent<E, 128, true>(tmA, tmB, p, splits, stream, cs);               \
  } while (0)
  if (epi == EPI_BF16) CCD_CONV_LAUNCH(EPI_BF16);
  CCD_CONV_LAUNCH(EPI_F32);
#undef CCD_CONV_LAUNCH
}

// debug / A-B switch: key 0 = GEMM variant (1 persistent, 0 one-tile-per-CTA); key 1 = epilogue of full tiles
// (1 per-shape choice [default], 0 shared-memory transpose, 2 transpose-free thread-per-row wherever alignment allows);
// key 2 = programmatic dependent launch of the GEMM / LayerNorm / attention kernels (0 off [default], 1 on)
extern "C" int ccd_set_option(int key, int value) {
  if (key == 0) { g_gemm_variant = value ? 1 : 0; return CCD_OK; }
  if (key == 1) { g_gemm_epilogue = (value < 0 || value > 2) ? 1 : value; return CCD_OK; }
  if (key == 2) { pdl_enabled() = value ? 1 : 0; return CCD_OK; }
  return CCD_ERR_ARG;
}

#if CCD_GEMM_TRACE
// diagnostic builds only (tools/gemm_trace.py): buf = device buffer of (11 * 4096 + 11) uint64 (zero-filled), NULL = off
extern "C" int ccd_debug_gemm_trace(unsigned long long* buf, unsigned int cap_unused) {
  (void)cap_unused;
  CCD_CUDA_CHECK(cudaMemcpyToSymbol(ccd::g_gemm_trace_buf, &buf, sizeof(buf)));
  return CCD_OK;
}
#endif
