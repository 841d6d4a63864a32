// Micro-benchmark: TMEM -> register read bandwidth (tcgen05.ld.32x32b.x32) per SM, 4 and 8 warps, 1..4 loads in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

template <int INFLIGHT>
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[INFLIGHT][32];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) ld32(base + ((i * INFLIGHT + j) * 32) % 512, a[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) acc ^= a[j][0] ^ a[j][31];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  const int iters = 2000;
  for (int threads : {128, 256}) {
    for (int inflight : {1, 2, 4}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (inflight == 1) k<1><<<1, threads>>>(iters, d, s);
        if (inflight == 2) k<2><<<1, threads>>>(iters, d, s);
        if (inflight == 4) k<4><<<1, threads>>>(iters, d, s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      const double bytes = (double)iters * inflight * (threads / 32) * 32 * 32 * 4;
      printf("{\"threads\": %d, \"loads_in_flight\": %d, \"cycles\": %lld, \"bytes_per_clk_per_sm\": %.1f}\n", threads, inflight, h, bytes / h);
    }
  }
  return 0;
}
