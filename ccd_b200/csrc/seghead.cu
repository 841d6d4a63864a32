// BatchNorm2d (training mode) + ReLU of the CCD segmentation head (Dino/modules/segmentor.py:37-95) on NHWC bf16
// activations [M = N*H*W, C]: HBM-bound row-streaming kernels.  The convolutions themselves run on the tcgen05
// implicit-GEMM path (gemm_umma.cu: ccd_conv_gemm).
//   forward : bn_stats (sum, sum of squares per channel)  -> [host glue: mean / rstd / running stats, SyncBN all-reduce]
//             bn_apply_relu: y = relu((x - mean) * rstd * gamma + beta)
//   backward: bn_bwd_reduce: dbeta = sum dz, dgamma = sum dz * xhat with dz = dy * (y > 0)   -> [SyncBN all-reduce]
//             bn_bwd_apply : dx = gamma * rstd * (dz - dbeta/M - xhat * dgamma/M)
#include "ccd_common.cuh"

namespace ccd {

constexpr int BN_ROWS_PER_BLOCK = 128;

// x [M, C] (ld = ldx), C % 8 == 0, C <= 256.  sums[0:C] += sum x, sums[C:2C] += sum x^2   (zero-filled by the caller)
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ sums, int M, int C) {
  __shared__ float acc[2][256];
  const int cg = threadIdx.x;                   // column group of 8 (blockDim = (C/8, 256/(C/8)))
  const int ty = threadIdx.y, ny = blockDim.y;
  for (int i = ty * blockDim.x + cg; i < 512; i += blockDim.x * blockDim.y) (&acc[0][0])[i] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * BN_ROWS_PER_BLOCK;
  const int r1 = min(M, r0 + BN_ROWS_PER_BLOCK);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  for (int r = r0 + ty; r < r1; r += ny) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8);
    const float f[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y), bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
  }
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

// y[r, yoff + c] = relu((x[r,c] - mean[c]) * rstd[c] * gamma[c] + beta[c])
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y, int ldy, int M, int C) {
  const int cgs = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * cgs) return;
  const int cg = (int)(idx % cgs);
  const size_t r = idx / cgs;
  const uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + cg * 8);
  const float f[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y), bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    const float sc = __ldg(rstd + c) * __ldg(gamma + c);
    o[j] = fmaxf(fmaf(f[j] - __ldg(mean + c), sc, __ldg(beta + c)), 0.f);
  }
  *reinterpret_cast<uint4*>(y + r * ldy + cg * 8) =
      make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// sums[0:C] += sum dz, sums[C:2C] += sum dz * xhat,   dz = dy * [xhat*gamma + beta > 0]
template <bool DY_F32>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const void* __restrict__ dy_, int lddy, const bf16* __restrict__ x, int ldx,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ sums, int M, int C) {
  __shared__ float acc[2][256];
  const int cg = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  for (int i = ty * blockDim.x + cg; i < 512; i += blockDim.x * blockDim.y) (&acc[0][0])[i] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * BN_ROWS_PER_BLOCK;
  const int r1 = min(M, r0 + BN_ROWS_PER_BLOCK);
  float s[8], q[8], mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    s[j] = 0.f; q[j] = 0.f; mu[j] = mean[c]; rs[j] = rstd[c]; ga[j] = gamma[c]; be[j] = beta[c];
  }
  for (int r = r0 + ty; r < r1; r += ny) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + (size_t)r * ldx + cg * 8);
    const float f[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y), bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
    float d[8];
    if constexpr (DY_F32) {
      const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)r * lddy + cg * 8);
      const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)r * lddy + cg * 8 + 4);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    } else {
      const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(dy_) + (size_t)r * lddy + cg * 8);
      d[0] = bf16lo(w.x); d[1] = bf16hi(w.x); d[2] = bf16lo(w.y); d[3] = bf16hi(w.y);
      d[4] = bf16lo(w.z); d[5] = bf16hi(w.z); d[6] = bf16lo(w.w); d[7] = bf16hi(w.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu[j]) * rs[j];
      const float dz = (fmaf(xh, ga[j], be[j]) > 0.f) ? d[j] : 0.f;
      s[j] += dz;
      q[j] += dz * xh;
    }
  }
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

// dx = gamma * rstd * (dz - dbeta * inv_m - xhat * dgamma * inv_m)
template <bool DY_F32>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const void* __restrict__ dy_, int lddy, const bf16* __restrict__ x, int ldx,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ sums, float inv_m, bf16* __restrict__ dx, int lddx,
                                                           int M, int C) {
  const int cgs = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * cgs) return;
  const int cg = (int)(idx % cgs);
  const size_t r = idx / cgs;
  const uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + cg * 8);
  const float f[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y), bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
  float d[8];
  if constexpr (DY_F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + r * lddy + cg * 8);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + r * lddy + cg * 8 + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
    const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(dy_) + r * lddy + cg * 8);
    d[0] = bf16lo(w.x); d[1] = bf16hi(w.x); d[2] = bf16lo(w.y); d[3] = bf16hi(w.y);
    d[4] = bf16lo(w.z); d[5] = bf16hi(w.z); d[6] = bf16lo(w.w); d[7] = bf16hi(w.w);
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    const float rs = __ldg(rstd + c), ga = __ldg(gamma + c);
    const float xh = (f[j] - __ldg(mean + c)) * rs;
    const float dz = (fmaf(xh, ga, __ldg(beta + c)) > 0.f) ? d[j] : 0.f;
    o[j] = ga * rs * (dz - __ldg(sums + c) * inv_m - xh * __ldg(sums + C + c) * inv_m);
  }
  *reinterpret_cast<uint4*>(dx + r * lddx + cg * 8) =
      make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
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
  if (!x || !y || !mean || !rstd || !gamma || !beta || M <= 0 || (C & 7)) return CCD_ERR_ARG;
  const size_t total = (size_t)M * (C / 8);
  bn_apply_relu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, mean, rstd, gamma, beta,
                                                                                          (bf16*)y, ldy, M, C);
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
  if (!dy || !x || !sums || !dx || M <= 0 || (C & 7)) return CCD_ERR_ARG;
  const size_t total = (size_t)M * (C / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dy_is_f32)
    bn_bwd_apply_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta, sums, inv_m,
                                                                        (bf16*)dx, lddx, M, C);
  else
    bn_bwd_apply_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma, beta, sums, inv_m,
                                                                         (bf16*)dx, lddx, M, C);
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
