// Micro-benchmark: the GEMM main loop of csrc/gemm_umma.cu WITHOUT its epilogue (accumulators are dropped), at the real
// problem shapes of the ViT-Small step, to find what caps the tcgen05 rate: operand majorness (K-major: one TMA box per
// operand and k-block; MN-major: 64x64 boxes), ring depth, one CTA per tile (cta_group::1, 128 x BN) versus CTA pairs
// (cta_group::2, 256 x BN, each CTA stages its 128 rows of A and HALF of B), resident operands versus TMA streaming.
// Items are scheduled exactly like the persistent kernel: item -> (k-slice z, tile), tile -> (m-block, n-block), n fastest.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../ccd_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "ccd_common.cuh"
#include "tmap.cuh"

using namespace ccd;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// TMA load whose completion bytes are signalled on an mbarrier given as a shared::cluster address (the pair leader's)
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct Prob {
  int M, N, K, a_mn, b_mn, splits, use_tma;
  int n_tiles_n, n_tiles, n_items, kb_per_split;
};

template <int CG, int BN, int STAGES>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Prob p, long long* cycles) {
  constexpr int A_BYTES = 128 * 64 * 2;
  constexpr int B_COLS = BN / CG;                 // columns of the N tile this CTA keeps in shared memory
  constexpr int B_BYTES = B_COLS * 64 * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;     // CTA or CTA pair
  const int kb_total = p.K / 64;

  for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);                 // one arrive.expect_tx per producer warp (of the leader CTA)
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    if (CG == 1) {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if ((warp == 0 || warp == 2) && lane == 0 && p.use_tma) {
    const bool is_a = (warp == 0);
    uint32_t kbc = 0;
    for (int item = unit; item < p.n_items; item += n_units) {
      const int z = item / p.n_tiles, tile = item - z * p.n_tiles;
      const int m0 = (tile / p.n_tiles_n) * (128 * CG) + (int)rank * 128;
      const int n0 = (tile % p.n_tiles_n) * BN + (int)rank * B_COLS;
      const int kb_begin = z * p.kb_per_split;
      const int nkb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
      for (int i = 0; i < nkb; ++i, ++kbc) {
        const int s = kbc % STAGES;
        const uint32_t ph = (kbc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = (kb_begin + i) * 64;
        if (CG == 1) {
          if (is_a) {
            mbar_arrive_expect_tx(&full_bar[s], A_BYTES);
            if (!p.a_mn) tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
            else { tma_load_2d(sa, &tmA, &full_bar[s], m0, k0); tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0); }
          } else {
            mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
            if (!p.b_mn) tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
            else {
#pragma unroll
              for (int j = 0; j < B_COLS / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
            }
          }
        } else {
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[s]), 0);
          if (is_a) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * A_BYTES);   // both CTAs' bytes land on the leader's barrier
            if (!p.a_mn) tma_load_2d_cg2(sa, &tmA, bar, k0, m0);
            else { tma_load_2d_cg2(sa, &tmA, bar, m0, k0); tma_load_2d_cg2(sa + 8192, &tmA, bar, m0 + 64, k0); }
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * B_BYTES);
            if (!p.b_mn) tma_load_2d_cg2(sb, &tmB, bar, k0, n0);
            else {
#pragma unroll
              for (int j = 0; j < B_COLS / 64; ++j) tma_load_2d_cg2(sb + j * 8192, &tmB, bar, n0 + 64 * j, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * CG, BN, p.a_mn, p.b_mn);
    const long long t0 = clock64();
    uint32_t kbc = 0;
    int it = 0;
    for (int item = unit; item < p.n_items; item += n_units, ++it) {
      const int z = item / p.n_tiles;
      const int kb_begin = z * p.kb_per_split;
      const int nkb = min(kb_total, kb_begin + p.kb_per_split) - kb_begin;
      const uint32_t d_tmem = tmem_base + (it & 1) * BN;
      for (int i = 0; i < nkb; ++i, ++kbc) {
        const int s = kbc % STAGES;
        const uint32_t ph = (kbc / STAGES) & 1;
        if (p.use_tma) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
        }
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = p.a_mn ? umma_smem_desc_sw128(a_addr + k * 2048, 8192, 1024) : umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = p.b_mn ? umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024) : umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          if (CG == 1) umma_ss(d_tmem, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else         umma_ss_cg2(d_tmem, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        if (CG == 1) umma_commit(&empty_bar[s]);
        else         umma_commit_cg2(&empty_bar[s], 3);
      }
    }
    if (CG == 1) umma_commit(done_bar);
    else         umma_commit_cg2(done_bar, 3);
    mbar_wait(done_bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  // the peer CTA must stay resident (its shared memory / TMEM are operands of the pair's MMAs) until the leader is done
  if (CG == 2 && rank == 1 && threadIdx.x == 32) mbar_wait(done_bar, 0);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static void *g_A, *g_B;          // two 512 MB operand pools
static long long* g_cycles;
static int g_sms = 148;

template <int CG, int BN, int STAGES>
static void run(const char* name, int M, int N, int K, int a_mn, int b_mn, int splits, int use_tma) {
  constexpr int B_COLS = BN / CG;
  constexpr int STAGE_BYTES = 128 * 64 * 2 + B_COLS * 64 * 2;
  const int smem = STAGES * STAGE_BYTES + 256 + 1024;
  if (b_mn && (B_COLS % 64)) { printf("{\"name\": \"%s\", \"skip\": \"MN-major B needs 64-column boxes\"}\n", name); return; }
  CUtensorMap tmA, tmB;
  bool ok = a_mn ? get_tmap_bf16_2d(&tmA, g_A, K, M, M, 64, 64) : get_tmap_bf16_2d(&tmA, g_A, M, K, K, 128, 64);
  ok = ok && (b_mn ? get_tmap_bf16_2d(&tmB, g_B, K, N, N, 64, 64) : get_tmap_bf16_2d(&tmB, g_B, N, K, K, B_COLS, 64));
  if (!ok) { printf("tensor map creation failed for %s\n", name); exit(1); }
  Prob p{M, N, K, a_mn, b_mn, splits, use_tma, 0, 0, 0, 0};
  const int tile_m = 128 * CG;
  p.n_tiles_n = (N + BN - 1) / BN;
  p.n_tiles = p.n_tiles_n * ((M + tile_m - 1) / tile_m);
  const int kb_total = K / 64;
  p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.n_items = p.n_tiles * p.splits;
  cudaFuncSetAttribute(mma_rate_kernel<CG, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int units = g_sms / CG;
  if (p.n_items < units) units = p.n_items;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  float best = 1e30f;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemsetAsync(g_cycles, 0, sizeof(long long) * 256);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, BN, STAGES>, tmA, tmB, p, g_cycles);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
      printf("{\"name\": \"%s\", \"cg\": %d, \"bn\": %d, \"error\": \"%s / %s\"}\n", name, CG, BN, cudaGetErrorString(e), cudaGetErrorString(e2));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long h[256];
  cudaMemcpy(h, g_cycles, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 256; ++i) mx = h[i] > mx ? h[i] : mx;
  const double useful = 2.0 * M * N * (double)K;
  const double issued = 2.0 * (double)p.n_tiles * tile_m * BN * (double)K;   // including padded tile area
  printf("{\"name\": \"%s\", \"cg\": %d, \"bn\": %d, \"stages\": %d, \"tma\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"a_mn\": %d, \"b_mn\": %d, "
         "\"splits\": %d, \"items\": %d, \"ctas\": %d, \"us\": %.1f, \"useful_tflops\": %.0f, \"issued_tflops\": %.0f, \"max_cycles\": %lld, "
         "\"ghz\": %.2f}\n",
         name, CG, BN, STAGES, use_tma, M, N, K, a_mn, b_mn, p.splits, p.n_items, units * CG, best * 1e3, useful / (best * 1e-3) / 1e12,
         issued / (best * 1e-3) / 1e12, mx, (double)mx / (best * 1e-3) / 1e9);
  fflush(stdout);
}

static int splits_for(int M, int N, int K, int tile_m, int bn, int units) {
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + bn - 1) / bn);
  int s = (2 * units + tiles - 1) / tiles;
  return s < 1 ? 1 : (s > K / 64 ? K / 64 : s);
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&g_A, (size_t)512 << 20);
  cudaMalloc(&g_B, (size_t)512 << 20);
  cudaMalloc(&g_cycles, sizeof(long long) * 256);
  cudaMemset(g_A, 0, (size_t)512 << 20);
  cudaMemset(g_B, 0, (size_t)512 << 20);
  const int T = 131072, E = 384;
  // --- resident operands (no TMA): the MMA issue ceiling
  run<1, 256, 4>("resident", T, 1536, 384, 0, 0, 1, 0);
  run<2, 256, 4>("resident", T, 1536, 384, 0, 0, 1, 0);
  // --- forward, K-major operands, K = 384 (fc1: N = 1536; qkv: N = 1152) and K = 1536 (fc2: N = 384)
  run<1, 256, 4>("fc1_fwd", T, 4 * E, E, 0, 0, 1, 1);
  run<2, 256, 4>("fc1_fwd", T, 4 * E, E, 0, 0, 1, 1);
  run<2, 256, 6>("fc1_fwd", T, 4 * E, E, 0, 0, 1, 1);
  run<1, 192, 4>("qkv_fwd", T, 3 * E, E, 0, 0, 1, 1);
  run<2, 192, 6>("qkv_fwd", T, 3 * E, E, 0, 0, 1, 1);
  run<2, 128, 8>("qkv_fwd", T, 3 * E, E, 0, 0, 1, 1);
  run<1, 192, 4>("fc2_fwd", T, E, 4 * E, 0, 0, 1, 1);
  run<2, 192, 6>("fc2_fwd", T, E, 4 * E, 0, 0, 1, 1);
  run<2, 128, 8>("fc2_fwd", T, E, 4 * E, 0, 0, 1, 1);
  // --- dgrad: B (the weight) MN-major
  run<1, 192, 4>("fc1_dgrad", T, E, 4 * E, 0, 1, 1, 1);
  run<2, 128, 8>("fc1_dgrad", T, E, 4 * E, 0, 1, 1, 1);
  run<2, 192, 6>("fc1_dgrad_kmajor_wT", T, E, 4 * E, 0, 0, 1, 1);
  run<1, 256, 4>("fc2_dgrad", T, 4 * E, E, 0, 1, 1, 1);
  run<2, 256, 6>("fc2_dgrad", T, 4 * E, E, 0, 1, 1, 1);
  // --- wgrad: both operands MN-major, split-K
  run<1, 256, 4>("fc2_wgrad", E, 4 * E, T, 1, 1, splits_for(E, 4 * E, T, 128, 256, 148), 1);
  run<2, 256, 6>("fc2_wgrad_swapped", 4 * E, E, T, 1, 1, splits_for(4 * E, E, T, 256, 256, 74), 1);
  run<2, 128, 8>("fc2_wgrad_swapped", 4 * E, E, T, 1, 1, splits_for(4 * E, E, T, 256, 128, 74), 1);
  run<1, 192, 4>("fc1_wgrad", 4 * E, E, T, 1, 1, splits_for(4 * E, E, T, 128, 192, 148), 1);
  run<2, 128, 8>("fc1_wgrad", 4 * E, E, T, 1, 1, splits_for(4 * E, E, T, 256, 128, 74), 1);
  run<2, 256, 6>("fc2_wgrad", E, 4 * E, T, 1, 1, splits_for(E, 4 * E, T, 256, 256, 74), 1);
  run<1, 192, 4>("qkv_wgrad", 3 * E, E, T, 1, 1, splits_for(3 * E, E, T, 128, 192, 148), 1);
  run<2, 128, 8>("qkv_wgrad", 3 * E, E, T, 1, 1, splits_for(3 * E, E, T, 256, 128, 74), 1);
  run<1, 128, 6>("proj_wgrad", E, E, T, 1, 1, splits_for(E, E, T, 128, 128, 148), 1);
  run<2, 128, 8>("proj_wgrad", E, E, T, 1, 1, splits_for(E, E, T, 256, 128, 74), 1);
  return 0;
}
