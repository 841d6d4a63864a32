// BatchNorm2d (training mode) + ReLU of the CCD segmentation head (Dino/modules/segmentor.py:37-95) on NHWC bf16
// activations [M = N*H*W, C]: HBM-bound row-streaming kernels.  The convolutions themselves run on the tcgen05
// implicit-GEMM path (gemm_umma.cu: ccd_conv_gemm).
//   forward : bn_stats (sum, sum of squares per channel)  -> [host glue: mean / rstd / running stats, SyncBN all-reduce]
//             bn_apply_relu: y = relu((x - mean) * rstd * gamma + beta)
//   backward: bn_bwd_reduce: dbeta = sum dz, dgamma = sum dz * xhat with dz = dy * (y > 0)   -> [SyncBN all-reduce]
//             bn_bwd_apply : dx = gamma * rstd * (dz - dbeta/M - xhat * dgamma/M)
#include "ccd_common.cuh"

namespace ccd {

// All four streaming kernels share one shape: blockDim = (C/8 column groups of 8 channels, 256/(C/8) row lanes); a
// thread keeps the per-channel constants of its 8 channels in registers and walks the rows of its block four at a time
// (four independent 16-byte loads in flight per operand).  First version: one thread per 16-byte group with ~40 __ldg of
// per-channel constants each (issue bound, 3.5x off the HBM floor in the step profile) and 128-row blocks (4 M global
// atomics per statistics pass).
constexpr int BN_ROWS_PER_BLOCK = 1024;   // statistics kernels: rows per block (global atomics: 2 C per block)
constexpr int BN_APPLY_ROWS = 256;        // apply kernels: rows per block

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
template <bool F32>
__device__ __forceinline__ void load_dy8(const void* dy_, size_t off, float (&d)[8]) {
  if constexpr (F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + off);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + off + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
    unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(dy_) + off), d);
  }
}
// block-level reduction of the per-thread partial sums: shared-memory atomics, then one global atomic per channel
__device__ __forceinline__ void bn_block_reduce(float (&s)[8], float (&q)[8], float* __restrict__ sums, int C) {
  __shared__ float acc[2][256];
  const int cg = threadIdx.x, ty = threadIdx.y;
  for (int i = ty * blockDim.x + cg; i < 512; i += blockDim.x * blockDim.y) (&acc[0][0])[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&acc[0][cg * 8 + j], s[j]);
    atomicAdd(&acc[1][cg * 8 + j], q[j]);
  }
  __syncthreads();
  for (int i = ty * blockDim.x + cg; i < C; i += blockDim.x * blockDim.y) {
    atomicAdd(sums + i, acc[0][i]);
    atomicAdd(sums + C + i, acc[1][i]);
  }
}

// x [M, C] (ld = ldx), C % 8 == 0, C <= 256.  sums[0:C] += sum x, sums[C:2C] += sum x^2   (zero-filled by the caller)
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ sums, int M, int C) {
  const int cg = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const int r0 = blockIdx.x * BN_ROWS_PER_BLOCK;
  const int r1 = min(M, r0 + BN_ROWS_PER_BLOCK);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  int r = r0 + ty;
  for (; r + 3 * ny < r1; r += 4 * ny) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(x + (size_t)(r + u * ny) * ldx + cg * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    }
  }
  for (; r < r1; r += ny) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
  }
  bn_block_reduce(s, q, sums, C);
}

// y[r, c] = relu((x[r,c] - mean[c]) * rstd[c] * gamma[c] + beta[c]) = relu(x * sc + sh)
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y, int ldy, int M, int C) {
  const int cg = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    sc[j] = rstd[c] * gamma[c];
    sh[j] = fmaf(-mean[c], sc[j], beta[c]);
  }
  const int r0 = blockIdx.x * BN_APPLY_ROWS;
  const int r1 = min(M, r0 + BN_APPLY_ROWS);
  auto one = [&](const uint4& v, int r) {
    float f[8], o[8];
    unpack8(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
    *reinterpret_cast<uint4*>(y + (size_t)r * ldy + cg * 8) =
        make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  };
  int r = r0 + ty;
  for (; r + 3 * ny < r1; r += 4 * ny) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(x + (size_t)(r + u * ny) * ldx + cg * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u) one(v[u], r + u * ny);
  }
  for (; r < r1; r += ny) one(*reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8), r);
}

// sums[0:C] += sum dz, sums[C:2C] += sum dz * xhat,   dz = dy * [xhat*gamma + beta > 0]
template <bool DY_F32>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const void* __restrict__ dy_, int lddy, const bf16* __restrict__ x, int ldx,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ sums, int M, int C) {
  const int cg = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const int r0 = blockIdx.x * BN_ROWS_PER_BLOCK;
  const int r1 = min(M, r0 + BN_ROWS_PER_BLOCK);
  // xhat = x * rs - mu * rs ;  y = xhat * gamma + beta = x * ysc + ysh
  float s[8], q[8], rs[8], mrs[8], ysc[8], ysh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    s[j] = 0.f; q[j] = 0.f;
    rs[j] = rstd[c];
    mrs[j] = -mean[c] * rs[j];
    ysc[j] = rs[j] * gamma[c];
    ysh[j] = fmaf(-mean[c], ysc[j], beta[c]);          // same expression as the forward: identical ReLU mask
  }
  auto one = [&](const uint4& v, const float (&d)[8]) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = (fmaf(f[j], ysc[j], ysh[j]) > 0.f) ? d[j] : 0.f;
      s[j] += dz;
      q[j] = fmaf(dz, fmaf(f[j], rs[j], mrs[j]), q[j]);
    }
  };
  int r = r0 + ty;
  for (; r + ny < r1; r += 2 * ny) {
    uint4 v[2];
    float d[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      v[u] = *reinterpret_cast<const uint4*>(x + (size_t)(r + u * ny) * ldx + cg * 8);
      load_dy8<DY_F32>(dy_, (size_t)(r + u * ny) * lddy + cg * 8, d[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) one(v[u], d[u]);
  }
  for (; r < r1; r += ny) {
    float d[8];
    load_dy8<DY_F32>(dy_, (size_t)r * lddy + cg * 8, d);
    one(*reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8), d);
  }
  bn_block_reduce(s, q, sums, C);
}

// dx = gamma * rstd * (dz - dbeta * inv_m - xhat * dgamma * inv_m)
template <bool DY_F32>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const void* __restrict__ dy_, int lddy, const bf16* __restrict__ x, int ldx,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ sums, float inv_m, bf16* __restrict__ dx, int lddx,
                                                           int M, int C) {
  const int cg = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  // dx = a * dz - a * (s1 + xhat * s2) ,  xhat = x * rs + mrs ,  a = gamma * rs  ->  dx = a * dz + x * k1 + k0
  float a[8], k0[8], k1[8], ysc[8], ysh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    const float rs = rstd[c], ga = gamma[c];
    const float mrs = -mean[c] * rs;
    const float s1 = sums[c] * inv_m, s2 = sums[C + c] * inv_m;
    a[j] = ga * rs;
    k1[j] = -a[j] * s2 * rs;
    k0[j] = -a[j] * (s1 + mrs * s2);
    ysc[j] = rs * ga;
    ysh[j] = fmaf(-mean[c], ysc[j], beta[c]);          // same expression as the forward: identical ReLU mask
  }
  const int r0 = blockIdx.x * BN_APPLY_ROWS;
  const int r1 = min(M, r0 + BN_APPLY_ROWS);
  auto one = [&](const uint4& v, const float (&d)[8], int r) {
    float f[8], o[8];
    unpack8(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = (fmaf(f[j], ysc[j], ysh[j]) > 0.f) ? d[j] : 0.f;
      o[j] = fmaf(a[j], dz, fmaf(f[j], k1[j], k0[j]));
    }
    *reinterpret_cast<uint4*>(dx + (size_t)r * lddx + cg * 8) =
        make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  };
  int r = r0 + ty;
  for (; r + ny < r1; r += 2 * ny) {
    uint4 v[2];
    float d[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      v[u] = *reinterpret_cast<const uint4*>(x + (size_t)(r + u * ny) * ldx + cg * 8);
      load_dy8<DY_F32>(dy_, (size_t)(r + u * ny) * lddy + cg * 8, d[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) one(v[u], d[u], r + u * ny);
  }
  for (; r < r1; r += ny) {
    float d[8];
    load_dy8<DY_F32>(dy_, (size_t)r * lddy + cg * 8, d);
    one(*reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8), d, r);
  }
}

// mean / rstd from the (possibly all-reduced) sums; running statistics with momentum (unbiased variance), as
// nn.BatchNorm2d / nn.SyncBatchNorm do in training mode.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float count, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - m * m, 0.f);
  mean[c] = m;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
  }
}


// ---------------------------------------------------------------------------------------------------------
// cls: Conv2d(128 -> 2, 3x3, pad 1) on the [n,32,128,128] bf16 feature map (segmentor.py:88,94).  Two output channels
// cannot feed a tcgen05 tile (a 128-wide UMMA tile would waste 98 % of its MACs).  First version (kept as variant 0): forward,
// data gradient and weight gradient as CUDA-core kernels, one CTA per image row with the three input rows staged in shared memory
// (pixel pitch 272 B = 256 B + 16 B pad: the per-pixel channel vectors of neighbouring pixels start in different banks).
// ---------------------------------------------------------------------------------------------------------
constexpr int CLS_W = 128, CLS_H = 32, CLS_C = 128;
constexpr int CLS_PITCH = 272;                          // bytes per pixel in smem
constexpr int CLS_ROW_BYTES = (CLS_W + 2) * CLS_PITCH;  // zero pixel on each side
constexpr int CLS_ROWS_BYTES = 3 * CLS_ROW_BYTES;       // 106080
constexpr int CLS_WS = CLS_C + 4;

// stage rows y-1, y, y+1 of image n (zeros outside the image) ; all threads of the CTA participate.
// Asynchronous 16-byte copies (cp.async, zero-fill for rows outside the image): every thread has its 24 copies in flight
// at once; the first version's load -> store loop exposed one global-memory latency per iteration (~7 us per image row).
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cls_stage_rows(uint8_t* rows, const bf16* __restrict__ u2, int n, int y) {
  for (int i = threadIdx.x; i < 3 * 2 * (CLS_PITCH / 16); i += blockDim.x) {        // left / right border pixels
    const int r = i / (2 * (CLS_PITCH / 16)), rem = i % (2 * (CLS_PITCH / 16));
    const int side = rem / (CLS_PITCH / 16), ch = rem % (CLS_PITCH / 16);
    *reinterpret_cast<uint4*>(rows + r * CLS_ROW_BYTES + (side ? (CLS_W + 1) : 0) * CLS_PITCH + ch * 16) = make_uint4(0, 0, 0, 0);
  }
  const uint32_t rows_s = smem_u32(rows);
  for (int i = threadIdx.x; i < 3 * CLS_W * 16; i += blockDim.x) {
    const int r = i / (CLS_W * 16), rem = i % (CLS_W * 16);
    const int px = rem >> 4, ch = rem & 15;
    const int yy = y + r - 1;
    const bool ok = yy >= 0 && yy < CLS_H;
    const bf16* src = ok ? u2 + (((size_t)n * CLS_H + yy) * CLS_W + px) * CLS_C + ch * 8 : u2;
    cp_async16(rows_s + r * CLS_ROW_BYTES + (px + 1) * CLS_PITCH + ch * 16, src, ok ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// logits[n,o,y,x] = bias[o] + sum_{ky,kx,c} u2[n, y+ky-1, x+kx-1, c] * w[o,c,ky,kx]        (NCHW fp32 output)
__global__ void __launch_bounds__(256) seg_cls_fwd_kernel(const bf16* __restrict__ u2, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ logits) {
  extern __shared__ __align__(16) uint8_t cls_smem[];
  uint8_t* rows = cls_smem;
  float* wsm = reinterpret_cast<float*>(cls_smem + CLS_ROWS_BYTES);      // [tap][o][c], class stride 132 floats (banks)
  const int n = blockIdx.x / CLS_H, y = blockIdx.x % CLS_H;
  for (int i = threadIdx.x; i < 9 * 2 * CLS_C; i += blockDim.x) {
    const int tap = i / (2 * CLS_C), o = (i / CLS_C) & 1, c = i % CLS_C;
    wsm[(tap * 2 + o) * CLS_WS + c] = w[(o * CLS_C + c) * 9 + tap];
  }
  cls_stage_rows(rows, u2, n, y);
  __syncthreads();
  const int x = threadIdx.x >> 1, o = threadIdx.x & 1;
  float acc = bias[o];
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    const uint8_t* px = rows + ky * CLS_ROW_BYTES + (x + kx) * CLS_PITCH;      // (x + kx - 1) + 1 border
    const float* wt = wsm + (tap * 2 + o) * CLS_WS;
#pragma unroll
    for (int cq = 0; cq < 16; ++cq) {
      const uint4 v = *reinterpret_cast<const uint4*>(px + cq * 16);
      const float4 w0 = *reinterpret_cast<const float4*>(wt + cq * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wt + cq * 8 + 4);
      acc += bf16lo(v.x) * w0.x + bf16hi(v.x) * w0.y + bf16lo(v.y) * w0.z + bf16hi(v.y) * w0.w + bf16lo(v.z) * w1.x +
             bf16hi(v.z) * w1.y + bf16lo(v.w) * w1.z + bf16hi(v.w) * w1.w;
    }
  }
  logits[(((size_t)n * 2 + o) * CLS_H + y) * CLS_W + x] = acc;
}

// du2[n,y,x,c] = sum_{ky,kx,o} dl[n,o, y-(ky-1), x-(kx-1)] * w[o,c,ky,kx]       (dl NCHW fp32, du2 NHWC bf16)
__global__ void __launch_bounds__(256) seg_cls_dgrad_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                                            bf16* __restrict__ du2) {
  __shared__ float dls[3][2][CLS_W + 2];
  const int n = blockIdx.x / CLS_H, y = blockIdx.x % CLS_H;
  for (int i = threadIdx.x; i < 3 * 2 * (CLS_W + 2); i += blockDim.x) {
    const int r = i / (2 * (CLS_W + 2)), o = (i / (CLS_W + 2)) & 1, xx = i % (CLS_W + 2) - 1;
    const int yy = y + r - 1;
    dls[r][o][xx + 1] = (yy >= 0 && yy < CLS_H && xx >= 0 && xx < CLS_W) ? dl[(((size_t)n * 2 + o) * CLS_H + yy) * CLS_W + xx] : 0.f;
  }
  // this thread's 4 channels of every tap / class: 72 weights in registers
  const int cg = threadIdx.x & 31, pl = threadIdx.x >> 5;          // channel group (4 ch), pixel lane (8 pixels per pass)
  float wr[9][2][4];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) wr[tap][o][j] = w[(o * CLS_C + cg * 4 + j) * 9 + tap];
  __syncthreads();
#pragma unroll 1
  for (int x = pl; x < CLS_W; x += 8) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const float d0 = dls[2 - ky][0][x + 2 - kx], d1 = dls[2 - ky][1][x + 2 - kx];     // row y+1-ky, col x+1-kx (+1 border)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = fmaf(d0, wr[tap][0][j], fmaf(d1, wr[tap][1][j], a[j]));
    }
    *reinterpret_cast<uint2*>(du2 + (((size_t)n * CLS_H + y) * CLS_W + x) * CLS_C + cg * 4) =
        make_uint2(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]));
  }
}

// dw[o,c,ky,kx] += sum_{n,y,x} dl[n,o,y,x] * u2[n, y+ky-1, x+kx-1, c]   ; persistent CTAs over image rows, dw zero-filled
__global__ void __launch_bounds__(256) seg_cls_wgrad_kernel(const bf16* __restrict__ u2, const float* __restrict__ dl,
                                                            float* __restrict__ dw, float* __restrict__ dbias, int n_rows_total) {
  extern __shared__ __align__(16) uint8_t cls_smem[];
  uint8_t* rows = cls_smem;
  float* dls = reinterpret_cast<float*>(cls_smem + CLS_ROWS_BYTES);      // [2][128]
  const int cp = threadIdx.x & 63, xq = threadIdx.x >> 6;                 // channel pair, pixel quarter
  float acc[9][2][2];                                                      // [tap][channel of the pair][class]
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t][0][0] = acc[t][0][1] = acc[t][1][0] = acc[t][1][1] = 0.f; }
  float bs0 = 0.f, bs1 = 0.f;                                              // cls.bias gradient (threads with cp == 0)
  for (int row = blockIdx.x; row < n_rows_total; row += gridDim.x) {
    const int n = row / CLS_H, y = row % CLS_H;
    __syncthreads();                                                       // previous row fully consumed
    cls_stage_rows(rows, u2, n, y);
    for (int i = threadIdx.x; i < 2 * CLS_W; i += blockDim.x)
      dls[i] = dl[(((size_t)n * 2 + (i >> 7)) * CLS_H + y) * CLS_W + (i & 127)];
    __syncthreads();
#pragma unroll 1
    for (int x = xq * 32; x < xq * 32 + 32; ++x) {
      const float d0 = dls[x], d1 = dls[CLS_W + x];
      bs0 += d0; bs1 += d1;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const uint32_t v = *reinterpret_cast<const uint32_t*>(rows + ky * CLS_ROW_BYTES + (x + kx) * CLS_PITCH + cp * 4);
        const float c0 = bf16lo(v), c1 = bf16hi(v);
        acc[tap][0][0] = fmaf(d0, c0, acc[tap][0][0]); acc[tap][0][1] = fmaf(d1, c0, acc[tap][0][1]);
        acc[tap][1][0] = fmaf(d0, c1, acc[tap][1][0]); acc[tap][1][1] = fmaf(d1, c1, acc[tap][1][1]);
      }
    }
  }
  // four pixel quarters hold partial sums of the same (o, c, tap): atomics into global (296 CTAs x 2304 addresses)
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int o = 0; o < 2; ++o) atomicAdd(dw + (o * CLS_C + cp * 2 + j) * 9 + tap, acc[tap][j][o]);
  if (cp == 0) { atomicAdd(dbias, bs0); atomicAdd(dbias + 1, bs1); }
}

// ---------------------------------------------------------------------------------------------------------
// cls on warp-level tensor cores (default).  The CUDA-core kernels above are shared-memory-bandwidth bound (one LDS.128
// per 8 FMAs forward: 660 us for 537 MB of input against an 82 us HBM floor at batch 256).  An m16n8k16 warp MMA has an
// 8-wide N: the two classes waste 3/4 of it, which costs nothing here.  fp32 operands (weights forward, the logits gradient
// backward) are split into bf16 hi + lo parts and multiplied in two MMAs, so the results keep fp32-operand accuracy
// (the activations are bf16 in memory already).  Forward and weight gradient walk blocks of 8 image rows with a 4-slot ring
// of row buffers: every input row is staged ONCE per block (cp.async, prefetched one row ahead) instead of three times.
// ---------------------------------------------------------------------------------------------------------
constexpr int CLS_PX = CLS_PITCH / 2;                   // pixel pitch in bf16 elements (136)
constexpr int CLS_RB = 8;                               // image rows per work item
constexpr int CLS_RING = 4 * CLS_ROW_BYTES;             // 141 440 B

__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16(v));
  lo = v - hi;
}
__device__ __forceinline__ void cls_zero_borders(uint8_t* ring) {
  for (int i = threadIdx.x; i < 4 * 2 * (CLS_PITCH / 16); i += blockDim.x) {
    const int slot = i / (2 * (CLS_PITCH / 16)), rem = i % (2 * (CLS_PITCH / 16));
    const int side = rem / (CLS_PITCH / 16), ch = rem % (CLS_PITCH / 16);
    *reinterpret_cast<uint4*>(ring + slot * CLS_ROW_BYTES + (side ? (CLS_W + 1) : 0) * CLS_PITCH + ch * 16) = make_uint4(0, 0, 0, 0);
  }
}
// image row yy of image n -> ring slot (yy + 1) & 3 (zeros when yy is outside the image); one commit group per call
__device__ __forceinline__ void cls_prefetch_row(uint8_t* ring, const bf16* __restrict__ u2, int n, int yy) {
  const uint32_t dst = smem_u32(ring + ((yy + 1) & 3) * CLS_ROW_BYTES);
  const bool ok = yy >= 0 && yy < CLS_H;
  for (int i = threadIdx.x; i < CLS_W * 16; i += blockDim.x) {
    const int px = i >> 4, ch = i & 15;
    const bf16* src = ok ? u2 + (((size_t)n * CLS_H + yy) * CLS_W + px) * CLS_C + ch * 8 : u2;
    cp_async16(dst + (px + 1) * CLS_PITCH + ch * 16, src, ok ? 16 : 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ const bf16* cls_row(const uint8_t* ring, int yy) {
  return reinterpret_cast<const bf16*>(ring + ((yy + 1) & 3) * CLS_ROW_BYTES);
}

// forward: warp w = pixels [16w, 16w+16) of the row; B fragments (weights, hi / lo) pre-formed in shared memory per (tap, k-step, lane)
__global__ void __launch_bounds__(256, 1) seg_cls_fwd_mma_kernel(const bf16* __restrict__ u2, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ logits, int n_items) {
  extern __shared__ __align__(16) uint8_t cls_smem[];
  uint8_t* ring = cls_smem;
  uint2* whi = reinterpret_cast<uint2*>(cls_smem + CLS_RING);          // [9 taps][8 k-steps][32 lanes]
  uint2* wlo = whi + 9 * 8 * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 9 * 8 * 32; i += blockDim.x) {
    const int l = i & 31, ks = (i >> 5) & 7, tap = i >> 8;
    const int o = l >> 2, c = ks * 16 + 2 * (l & 3);
    float h[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
    if (o < 2) {
      split_bf16(w[(o * CLS_C + c) * 9 + tap], h[0], lo[0]);
      split_bf16(w[(o * CLS_C + c + 1) * 9 + tap], h[1], lo[1]);
      split_bf16(w[(o * CLS_C + c + 8) * 9 + tap], h[2], lo[2]);
      split_bf16(w[(o * CLS_C + c + 9) * 9 + tap], h[3], lo[3]);
    }
    whi[i] = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
    wlo[i] = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
  }
  cls_zero_borders(ring);
  const float b0 = bias[0], b1 = bias[1];
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int n = item / (CLS_H / CLS_RB), y0 = (item % (CLS_H / CLS_RB)) * CLS_RB;
    __syncthreads();                                    // the previous item's rows are no longer read
    cls_prefetch_row(ring, u2, n, y0 - 1);
    cls_prefetch_row(ring, u2, n, y0);
    cls_prefetch_row(ring, u2, n, y0 + 1);
    for (int r = 0; r < CLS_RB; ++r) {
      const int y = y0 + r;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                  // rows y-1 .. y+1 visible; row y-2's slot is free
      if (r + 1 < CLS_RB) cls_prefetch_row(ring, u2, n, y + 2);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const bf16* rowp = cls_row(ring, y + ky - 1) + (warp * 16 + (lane & 15) + kx) * CLS_PX + (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t a[4];
          ldsm_x4(a, rowp + ks * 16);
          const uint2 bh = whi[(tap * 8 + ks) * 32 + lane], bl = wlo[(tap * 8 + ks) * 32 + lane];
          mma_bf16(acc, a, bh.x, bh.y);
          mma_bf16(acc, a, bl.x, bl.y);
        }
      }
      if ((lane & 3) == 0) {
        const int x = warp * 16 + (lane >> 2);
        float* out = logits + (((size_t)n * 2) * CLS_H + y) * CLS_W;
        out[x] = acc[0] + b0;
        out[(size_t)CLS_H * CLS_W + x] = acc[1] + b1;
        out[x + 8] = acc[2] + b0;
        out[(size_t)CLS_H * CLS_W + x + 8] = acc[3] + b1;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// weight gradient: warp w = channels [16w, 16w+16) for all 9 taps; A = (dl row, hi / lo) as a 16 x 16 tile whose rows 0, 1 are the classes
__global__ void __launch_bounds__(256, 1) seg_cls_wgrad_mma_kernel(const bf16* __restrict__ u2, const float* __restrict__ dl,
                                                                   float* __restrict__ dw, float* __restrict__ dbias, int n_items) {
  extern __shared__ __align__(16) uint8_t cls_smem[];
  uint8_t* ring = cls_smem;
  float* dls = reinterpret_cast<float*>(cls_smem + CLS_RING);          // [2 buffers][2 classes][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cls_zero_borders(ring);
  float acc[9][2][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][j][e] = 0.f;
  float bsum = 0.f;
  const int o_b = threadIdx.x >> 7, x_b = threadIdx.x & 127;           // dl element this thread stages / sums
  int it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int n = item / (CLS_H / CLS_RB), y0 = (item % (CLS_H / CLS_RB)) * CLS_RB;
    __syncthreads();
    cls_prefetch_row(ring, u2, n, y0 - 1);
    cls_prefetch_row(ring, u2, n, y0);
    cls_prefetch_row(ring, u2, n, y0 + 1);
    for (int r = 0; r < CLS_RB; ++r, ++it) {
      const int y = y0 + r;
      float* dcur = dls + (it & 1) * 2 * CLS_W;
      const float dv = dl[(((size_t)n * 2 + o_b) * CLS_H + y) * CLS_W + x_b];
      dcur[o_b * CLS_W + x_b] = dv;                     // double buffered: the other buffer may still be read by slower warps
      bsum += dv;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (r + 1 < CLS_RB) cls_prefetch_row(ring, u2, n, y + 2);
      uint32_t ah[8][2], al[8][2];                      // a0 / a2 of the 8 pixel k-steps (rows 8..15 of the tile are zero)
      const int o = lane >> 2;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        float h[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
        if (o < 2) {
          const float* d = dcur + o * CLS_W + ks * 16 + 2 * (lane & 3);
          split_bf16(d[0], h[0], lo[0]); split_bf16(d[1], h[1], lo[1]);
          split_bf16(d[8], h[2], lo[2]); split_bf16(d[9], h[3], lo[3]);
        }
        ah[ks][0] = pack_bf16x2(h[0], h[1]); ah[ks][1] = pack_bf16x2(h[2], h[3]);
        al[ks][0] = pack_bf16x2(lo[0], lo[1]); al[ks][1] = pack_bf16x2(lo[2], lo[3]);
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const int mi = lane >> 3;
        const bf16* rowp = cls_row(ring, y + ky - 1) + ((lane & 7) + (mi & 1) * 8 + kx) * CLS_PX + warp * 16 + (mi >> 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t b[4];
          ldsm_x4_t(b, rowp + ks * 16 * CLS_PX);        // B[k = pixel][n = channel]: two channel tiles
          const uint32_t a_h[4] = {ah[ks][0], 0u, ah[ks][1], 0u}, a_l[4] = {al[ks][0], 0u, al[ks][1], 0u};
          mma_bf16(acc[tap][0], a_h, b[0], b[1]);
          mma_bf16(acc[tap][0], a_l, b[0], b[1]);
          mma_bf16(acc[tap][1], a_h, b[2], b[3]);
          mma_bf16(acc[tap][1], a_l, b[2], b[3]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if ((lane >> 2) < 2) {
    const int o = lane >> 2;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = warp * 16 + j * 8 + 2 * (lane & 3);
        atomicAdd(dw + (o * CLS_C + c) * 9 + tap, acc[tap][j][0]);
        atomicAdd(dw + (o * CLS_C + c + 1) * 9 + tap, acc[tap][j][1]);
      }
  }
  bsum = warp_sum(bsum);
  if (lane == 0) atomicAdd(dbias + (warp >> 2), bsum);
}

// data gradient: du2[px, c] = sum_k A[px, k] B[k, c], k = (tap, class) (18 -> 32): warp w = pixels [16w, 16w+16), all 128 channels
__global__ void __launch_bounds__(256) seg_cls_dgrad_mma_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                                                bf16* __restrict__ du2) {
  __shared__ float dls[3][2][CLS_W + 2];
  __shared__ uint2 wfr[2][16][32];                       // B fragments [k-step][channel tile][lane], bf16 hi part
  __shared__ uint2 wfl[2][16][32];                       // lo part (w - hi)
  const int n = blockIdx.x / CLS_H, y = blockIdx.x % CLS_H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * 2 * (CLS_W + 2); i += blockDim.x) {
    const int r = i / (2 * (CLS_W + 2)), o = (i / (CLS_W + 2)) & 1, xx = i % (CLS_W + 2) - 1;
    const int yy = y + r - 1;
    dls[r][o][xx + 1] = (yy >= 0 && yy < CLS_H && xx >= 0 && xx < CLS_W) ? dl[(((size_t)n * 2 + o) * CLS_H + yy) * CLS_W + xx] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * 16 * 32; i += blockDim.x) {
    const int l = i & 31, nt = (i >> 5) & 15, ks = i >> 9;
    const int c = nt * 8 + (l >> 2), t0 = ks * 8 + (l & 3), t1 = t0 + 4;
    const float w00 = t0 < 9 ? w[(0 * CLS_C + c) * 9 + t0] : 0.f, w01 = t0 < 9 ? w[(1 * CLS_C + c) * 9 + t0] : 0.f;
    const float w10 = t1 < 9 ? w[(0 * CLS_C + c) * 9 + t1] : 0.f, w11 = t1 < 9 ? w[(1 * CLS_C + c) * 9 + t1] : 0.f;
    float h[4], lo[4];
    split_bf16(w00, h[0], lo[0]); split_bf16(w01, h[1], lo[1]); split_bf16(w10, h[2], lo[2]); split_bf16(w11, h[3], lo[3]);
    wfr[ks][nt][l] = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
    wfl[ks][nt][l] = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
  }
  __syncthreads();
  // A fragments: k pair (2 (lane & 3), +1) = (tap t, class 0 / 1); value = dl[class][y + 1 - ky][x + 1 - kx]
  uint32_t ah[2][4], al[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int t = ks * 8 + (lane & 3) + (e >> 1) * 4, px = warp * 16 + (lane >> 2) + (e & 1) * 8;
      float h0 = 0.f, l0 = 0.f, h1 = 0.f, l1 = 0.f;
      if (t < 9) {
        const int ky = t / 3, kx = t % 3;
        split_bf16(dls[2 - ky][0][px + 2 - kx], h0, l0);
        split_bf16(dls[2 - ky][1][px + 2 - kx], h1, l1);
      }
      ah[ks][e] = pack_bf16x2(h0, h1);
      al[ks][e] = pack_bf16x2(l0, l1);
    }
  bf16* orow = du2 + (((size_t)n * CLS_H + y) * CLS_W + warp * 16 + (lane >> 2)) * CLS_C + 2 * (lane & 3);
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint2 b = wfr[ks][nt][lane], bl = wfl[ks][nt][lane];
      mma_bf16(acc, ah[ks], b.x, b.y);
      mma_bf16(acc, al[ks], b.x, b.y);
      mma_bf16(acc, ah[ks], bl.x, bl.y);
    }
    *reinterpret_cast<uint32_t*>(orow + nt * 8) = pack_bf16x2(acc[0], acc[1]);
    *reinterpret_cast<uint32_t*>(orow + (size_t)8 * CLS_C + nt * 8) = pack_bf16x2(acc[2], acc[3]);
  }
}

static int g_cls_variant = 1;                           // 1 = warp-MMA kernels (default), 0 = CUDA-core kernels

}  // namespace ccd

using namespace ccd;

static inline dim3 bn_block(int C) {
  const int cgs = C / 8;
  return dim3(cgs, 256 / cgs > 0 ? 256 / cgs : 1);
}

extern "C" int ccd_bn_stats(const void* x, int ldx, float* sums_zeroed, int M, int C, void* stream) {
  if (!x || !sums_zeroed || M <= 0 || C <= 0 || (C & 7) || C > 256 || (256 % (C / 8))) return CCD_ERR_ARG;
  bn_stats_kernel<<<(M + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK, bn_block(C), 0, (cudaStream_t)stream>>>((const bf16*)x, ldx,
                                                                                                           sums_zeroed, M, C);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_bn_apply_relu(const void* x, int ldx, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                 void* y, int ldy, int M, int C, void* stream) {
  if (!x || !y || !mean || !rstd || !gamma || !beta || M <= 0 || (C & 7) || C > 256 || (256 % (C / 8))) return CCD_ERR_ARG;
  bn_apply_relu_kernel<<<(M + BN_APPLY_ROWS - 1) / BN_APPLY_ROWS, bn_block(C), 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, mean, rstd, gamma, beta, (bf16*)y, ldy, M, C);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_bn_bwd_reduce(const void* dy, int dy_is_f32, int lddy, const void* x, int ldx, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, float* sums_zeroed, int M, int C, void* stream) {
  if (!dy || !x || !sums_zeroed || M <= 0 || (C & 7) || C > 256 || (256 % (C / 8))) return CCD_ERR_ARG;
  const int blocks = (M + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK;
  if (dy_is_f32)
    bn_bwd_reduce_kernel<true><<<blocks, bn_block(C), 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta,
                                                                                 sums_zeroed, M, C);
  else
    bn_bwd_reduce_kernel<false><<<blocks, bn_block(C), 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta,
                                                                                  sums_zeroed, M, C);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_bn_bwd_apply(const void* dy, int dy_is_f32, int lddy, const void* x, int ldx, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, const float* sums, float inv_m, void* dx, int lddx, int M, int C,
                                void* stream) {
  if (!dy || !x || !sums || !dx || M <= 0 || (C & 7) || C > 256 || (256 % (C / 8))) return CCD_ERR_ARG;
  const unsigned blocks = (unsigned)((M + BN_APPLY_ROWS - 1) / BN_APPLY_ROWS);
  if (dy_is_f32)
    bn_bwd_apply_kernel<true><<<blocks, bn_block(C), 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta, sums,
                                                                                inv_m, (bf16*)dx, lddx, M, C);
  else
    bn_bwd_apply_kernel<false><<<blocks, bn_block(C), 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta, sums,
                                                                                 inv_m, (bf16*)dx, lddx, M, C);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_bn_finalize(const float* sums, float count, float eps, float momentum, float* mean, float* rstd,
                               float* running_mean, float* running_var, int C, void* stream) {
  if (!sums || !mean || !rstd || C <= 0 || count <= 0.f) return CCD_ERR_ARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, eps, momentum, mean, rstd, running_mean, running_var, C);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_seg_cls_fwd(const void* u2, const float* w, const float* bias, float* logits, int n_img, void* stream) {
  if (!u2 || !w || !bias || !logits || n_img <= 0) return CCD_ERR_ARG;
  const int smem = CLS_ROWS_BYTES + 9 * 2 * CLS_WS * 4;
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(seg_cls_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (g_cls_variant == 1) {
    const int smem2 = CLS_RING + 2 * 9 * 8 * 32 * 8;
    static bool attr2 = false;
    if (!attr2) {
      CCD_CUDA_CHECK(cudaFuncSetAttribute(seg_cls_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr2 = true;
    }
    const int items = n_img * (CLS_H / CLS_RB);
    seg_cls_fwd_mma_kernel<<<items < 148 ? items : 148, 256, smem2, (cudaStream_t)stream>>>((const bf16*)u2, w, bias, logits, items);
    CCD_LAUNCH_CHECK();
    return CCD_OK;
  }
  seg_cls_fwd_kernel<<<n_img * CLS_H, 256, smem, (cudaStream_t)stream>>>((const bf16*)u2, w, bias, logits);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_seg_cls_dgrad(const float* dl, const float* w, void* du2, int n_img, void* stream) {
  if (!dl || !w || !du2 || n_img <= 0) return CCD_ERR_ARG;
  if (g_cls_variant == 1) seg_cls_dgrad_mma_kernel<<<n_img * CLS_H, 256, 0, (cudaStream_t)stream>>>(dl, w, (bf16*)du2);
  else seg_cls_dgrad_kernel<<<n_img * CLS_H, 256, 0, (cudaStream_t)stream>>>(dl, w, (bf16*)du2);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_seg_cls_wgrad(const void* u2, const float* dl, float* dw_zeroed, float* dbias_zeroed, int n_img, void* stream) {
  if (!u2 || !dl || !dw_zeroed || !dbias_zeroed || n_img <= 0) return CCD_ERR_ARG;
  const int smem = CLS_ROWS_BYTES + 2 * CLS_W * 4;
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(seg_cls_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (g_cls_variant == 1) {
    const int smem2 = CLS_RING + 2 * 2 * CLS_W * 4;
    static bool attr2 = false;
    if (!attr2) {
      CCD_CUDA_CHECK(cudaFuncSetAttribute(seg_cls_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr2 = true;
    }
    const int items = n_img * (CLS_H / CLS_RB);
    seg_cls_wgrad_mma_kernel<<<items < 148 ? items : 148, 256, smem2, (cudaStream_t)stream>>>((const bf16*)u2, dl, dw_zeroed, dbias_zeroed, items);
    CCD_LAUNCH_CHECK();
    return CCD_OK;
  }
  const int rows = n_img * CLS_H;
  const int grid = rows < 296 ? rows : 296;
  seg_cls_wgrad_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)u2, dl, dw_zeroed, dbias_zeroed, rows);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

// A/B switch: 1 = warp-MMA classifier-convolution kernels (default), 0 = CUDA-core kernels
extern "C" int ccd_set_seg_cls_variant(int v) {
  g_cls_variant = v ? 1 : 0;
  return CCD_OK;
}
