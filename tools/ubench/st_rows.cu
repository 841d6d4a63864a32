// Micro-benchmark: global-store throughput of the GEMM epilogue's two possible register layouts for a bf16 [T, N] output.
//   mode 0  "transposed": a warp instruction writes 4 rows x 64 B (lane = (row & 3, 8-byte piece)) -- the layout the
//           epilogue reaches through its shared-memory transpose (STG.64, 4 cache lines per instruction)
//   mode 1  "thread = row", 2 x STG.256 per 32-column chunk (each lane owns 64 contiguous bytes of its own row: 32 lines
//           per instruction, one full 32 B sector per lane)
//   mode 2  "thread = row", 4 x STG.128
// Each warp walks 32-row x 32-column chunks of a [T, N] bf16 matrix like the epilogue warps do (tile 128 x 256).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o st_rows st_rows.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 1) st_rows_kernel(uint16_t* out, int T, int N, int n_tiles_n, int n_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, g = warp >> 2;                  // row quarter of the 128-row tile, column half
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = (tile / n_tiles_n) * 128 + q * 32, n0 = (tile % n_tiles_n) * 256 + g * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col = n0 + c * 32;
      if (MODE == 0) {
        const int sub_row = lane >> 3, sub = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint16_t* p = out + (size_t)(m0 + i * 4 + sub_row) * N + col + sub * 4;
          asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(tile + i), "r"(c) : "memory");
        }
      } else if (MODE == 1) {
        uint16_t* p = out + (size_t)(m0 + lane) * N + col;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + j * 16), "r"(tile), "r"(c), "r"(j), "r"(tile), "r"(c),
                       "r"(j), "r"(tile), "r"(c)
                       : "memory");
      } else {
        uint16_t* p = out + (size_t)(m0 + lane) * N + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p + j * 8), "r"(tile), "r"(c), "r"(j), "r"(tile) : "memory");
      }
    }
  }
}

int main() {
  const int T = 131072, N = 1536;
  uint16_t* d;
  cudaMalloc(&d, (size_t)T * N * 2);
  void* flush;
  cudaMalloc(&flush, (size_t)512 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n_tiles_n = N / 256, n_tiles = n_tiles_n * (T / 128);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemsetAsync(flush, rep, (size_t)512 << 20);
      cudaEventRecord(e0);
      if (mode == 0) st_rows_kernel<0><<<148, 256>>>(d, T, N, n_tiles_n, n_tiles);
      if (mode == 1) st_rows_kernel<1><<<148, 256>>>(d, T, N, n_tiles_n, n_tiles);
      if (mode == 2) st_rows_kernel<2><<<148, 256>>>(d, T, N, n_tiles_n, n_tiles);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("{\"mode\": %d, \"us\": %.1f, \"gbs\": %.0f, \"bytes_per_clk_per_sm_at_1p8ghz\": %.1f, \"err\": \"%s\"}\n", mode, best * 1e3,
           (double)T * N * 2 / (best * 1e-3) / 1e9, (double)T * N * 2 / (best * 1e-3) / 148 / 1.8e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
