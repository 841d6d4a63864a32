// Fused optimizer step of the CCD pretraining loop: per-parameter gradient clip + AdamW + teacher EMA + refresh of the
// bf16 GEMM operand copies, ONE pass over the parameter state (reference: the caller-side sequence
//   utils.clip_gradients            Dino/modules/utils.py:132-141   (one .item() host sync per parameter)
//   optimizer.step()  (torch AdamW) train.py:131-133,252
//   teacher EMA loop                train.py:264-272                (two tiny kernels per parameter)
// followed, in this repository, by the fp32 -> bf16 weight casts of the next forward).
// HBM-bound: 20 B read + 16 B written per parameter (+ 2 x 2 B bf16 copies for GEMM weights); algorithmic floor at
// 46.4 M parameters (ViT-Small student) ~ 1.7 GB.
#include "ccd_common.cuh"

namespace ccd {

// table row (int64 x 10): param, grad_ref, exp_avg, exp_avg_sq, sqnorm_ptr, ema_dst, param_bf16, ema_bf16, n_elems, flags
// grad_ref = tensor_index | (element_offset << 16): the gradient of tensor i lives at grad_ptrs[i] (a small device array
// refreshed every step), so the table itself survives gradient buffers moving between steps.
constexpr int OPT_COLS = 10;
enum { OPT_FLAG_WD = 1, OPT_FLAG_EMA_ONLY = 2 };

struct AdamWParams {
  float lr, beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, clip, ema_m;
};

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamWParams& a, float decay, float coef) {
  g *= coef;
  p *= decay;                                        // param.mul_(1 - lr * weight_decay)
  m = fmaf(a.beta1, m, (1.0f - a.beta1) * g);        // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, (1.0f - a.beta2) * g * g);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = __fsqrt_rn(v) * a.inv_sqrt_bc2 + a.eps;   // exact ops: the build uses --use_fast_math
  p -= a.step_size * __fdiv_rn(m, denom);                   // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
}

__device__ __forceinline__ const float* grad_of(const long long* __restrict__ grad_ptrs, long long ref) {
  return reinterpret_cast<const float*>(grad_ptrs[ref & 0xFFFF]) + (ref >> 16);
}

// *sqnorm[tensor] += sum(g^2) over the chunk (per-parameter L2 norms for the clip; sqnorm zero-filled by the caller)
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const long long* __restrict__ table,
                                                          const long long* __restrict__ grad_ptrs) {
  const long long* e = table + (size_t)blockIdx.x * OPT_COLS;
  float* sq = reinterpret_cast<float*>(e[4]);
  if (sq == nullptr || ((int)e[9] & OPT_FLAG_EMA_ONLY)) return;
  const float* g = grad_of(grad_ptrs, e[1]);
  const int n = (int)e[8];
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(g)[i];
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  }
  __shared__ float sm[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[w];
    atomicAdd(sq, t);
  }
}

__global__ void __launch_bounds__(256) fused_adamw_kernel(const long long* __restrict__ table,
                                                          const long long* __restrict__ grad_ptrs, const AdamWParams a) {
  const long long* e = table + (size_t)blockIdx.x * OPT_COLS;
  float* p = reinterpret_cast<float*>(e[0]);
  const bool ema_only_row = ((int)e[9] & OPT_FLAG_EMA_ONLY) != 0;
  const float* g = ema_only_row ? nullptr : grad_of(grad_ptrs, e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const float* sq = reinterpret_cast<const float*>(e[4]);
  float* t = reinterpret_cast<float*>(e[5]);
  bf16* pb = reinterpret_cast<bf16*>(e[6]);
  bf16* tb = reinterpret_cast<bf16*>(e[7]);
  const int n = (int)e[8];
  const int flags = (int)e[9];
  const bool ema_only = flags & OPT_FLAG_EMA_ONLY;
  const float decay = (flags & OPT_FLAG_WD) ? 1.0f - a.lr * a.weight_decay : 1.0f;
  float coef = 1.0f;
  if (sq != nullptr && a.clip > 0.f) {               // clip_coef = clip / (norm + 1e-6); applied only when < 1
    const float c = __fdiv_rn(a.clip, __fsqrt_rn(*sq) + 1e-6f);
    if (c < 1.0f) coef = c;
  }
  const float em = a.ema_m, em1 = 1.0f - a.ema_m;
  const uintptr_t align = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)t;
  const bool vec = ((align & 15) == 0) && (((uintptr_t)pb & 7) == 0) && (((uintptr_t)tb & 7) == 0);
  const int n4 = vec ? (n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 P = reinterpret_cast<const float4*>(p)[i];
    if (!ema_only) {
      const float4 G = reinterpret_cast<const float4*>(g)[i];
      float4 M = reinterpret_cast<const float4*>(m)[i];
      float4 V = reinterpret_cast<const float4*>(v)[i];
      adamw_elem(P.x, G.x, M.x, V.x, a, decay, coef);
      adamw_elem(P.y, G.y, M.y, V.y, a, decay, coef);
      adamw_elem(P.z, G.z, M.z, V.z, a, decay, coef);
      adamw_elem(P.w, G.w, M.w, V.w, a, decay, coef);
      reinterpret_cast<float4*>(p)[i] = P;
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
      if (pb != nullptr) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16x2(P.x, P.y), pack_bf16x2(P.z, P.w));
    }
    if (t != nullptr) {
      float4 Tt = reinterpret_cast<const float4*>(t)[i];
      Tt.x = em * Tt.x + em1 * P.x; Tt.y = em * Tt.y + em1 * P.y; Tt.z = em * Tt.z + em1 * P.z; Tt.w = em * Tt.w + em1 * P.w;
      reinterpret_cast<float4*>(t)[i] = Tt;
      if (tb != nullptr) reinterpret_cast<uint2*>(tb)[i] = make_uint2(pack_bf16x2(Tt.x, Tt.y), pack_bf16x2(Tt.z, Tt.w));
    }
  }
  for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
    float P = p[i];
    if (!ema_only) {
      float M = m[i], V = v[i];
      adamw_elem(P, g[i], M, V, a, decay, coef);
      p[i] = P; m[i] = M; v[i] = V;
      if (pb != nullptr) pb[i] = __float2bfloat16(P);
    }
    if (t != nullptr) {
      const float Tt = em * t[i] + em1 * P;
      t[i] = Tt;
      if (tb != nullptr) tb[i] = __float2bfloat16(Tt);
    }
  }
}

}  // namespace ccd

using namespace ccd;

// C ABI -- see include/ccd_b200.h
extern "C" int ccd_grad_sqnorm(const void* table_dev, const void* grad_ptrs_dev, int n_chunks, void* stream) {
  if (!table_dev || !grad_ptrs_dev || n_chunks <= 0) return CCD_ERR_ARG;
  grad_sqnorm_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const long long*)table_dev, (const long long*)grad_ptrs_dev);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_fused_adamw(const void* table_dev, const void* grad_ptrs_dev, int n_chunks, float lr, float beta1,
                               float beta2, float eps, float weight_decay, float bias_correction1, float bias_correction2,
                               float clip, float ema_m, void* stream) {
  if (!table_dev || !grad_ptrs_dev || n_chunks <= 0 || bias_correction1 <= 0.f || bias_correction2 <= 0.f) return CCD_ERR_ARG;
  AdamWParams a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.step_size = lr / bias_correction1;
  a.inv_sqrt_bc2 = 1.0f / sqrtf(bias_correction2);
  a.clip = clip; a.ema_m = ema_m;
  fused_adamw_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const long long*)table_dev, (const long long*)grad_ptrs_dev, a);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
