// Sharpened-softmax cross-view distillation loss (DINOLoss.forward, Dino/loss/Dino_loss.py:81-105) and the
// segmentation cross-entropy on already-softmaxed probabilities (Dino_loss.py:63-68,15-26), forward + backward.
// HBM-bound: one pass over the student row and its cross-view teacher row for the forward (online softmax of
// both + the q.z dot product), one pass for the backward writing bf16 dlogits.
//   student row i pairs with teacher row (i + R) mod 2R   (ncrops = 2: view-0 teacher x view-1 student and v.v.)
//   loss_i = logsumexp(z_i/ts) - sum_k q_k z_ik/ts ,  q = softmax((t - c)/tt)
//   L = 1/(2R) sum_i loss_i
#include "ccd_common.cuh"

namespace ccd {

constexpr int CE_THREADS = 512;

struct OnlineState {   // per-thread online softmax state of both rows
  float ms, ls;        // student: running max of z/ts, sum exp
  float mt, lt, dt;    // teacher: running max of (t-c)/tt, sum exp, sum exp * (z/ts)
};

__device__ __forceinline__ void online_merge(OnlineState& a, const OnlineState& b) {
  // states of threads that saw no element carry max = -inf: guard the (-inf) - (-inf) case
  const float ms = fmaxf(a.ms, b.ms);
  const float sa = (a.ms == -INFINITY) ? 0.f : __expf(a.ms - ms), sb = (b.ms == -INFINITY) ? 0.f : __expf(b.ms - ms);
  a.ls = a.ls * sa + b.ls * sb;
  a.ms = ms;
  const float mt = fmaxf(a.mt, b.mt);
  const float fa = (a.mt == -INFINITY) ? 0.f : __expf(a.mt - mt), fb = (b.mt == -INFINITY) ? 0.f : __expf(b.mt - mt);
  a.lt = a.lt * fa + b.lt * fb;
  a.dt = a.dt * fa + b.dt * fb;
  a.mt = mt;
}

__global__ void __launch_bounds__(CE_THREADS) dino_ce_fwd_kernel(const float* __restrict__ zs, const float* __restrict__ zt,
                                                                 const float* __restrict__ center, float inv_ts, float inv_tt,
                                                                 float* __restrict__ row_loss, float* __restrict__ stats,
                                                                 int R, int K) {
  const int i = blockIdx.x;
  const int j = (i + R) % (2 * R);
  const float4* ps = reinterpret_cast<const float4*>(zs + (size_t)i * K);
  const float4* pt = reinterpret_cast<const float4*>(zt + (size_t)j * K);
  const float4* pc = reinterpret_cast<const float4*>(center);
  OnlineState st{-INFINITY, 0.f, -INFINITY, 0.f, 0.f};
  for (int c = threadIdx.x; c < (K >> 2); c += CE_THREADS) {
    const float4 a = ps[c];
    const float4 t = pt[c];
    const float4 cc = __ldg(pc + c);
    const float z[4] = {a.x * inv_ts, a.y * inv_ts, a.z * inv_ts, a.w * inv_ts};
    const float u[4] = {(t.x - cc.x) * inv_tt, (t.y - cc.y) * inv_tt, (t.z - cc.z) * inv_tt, (t.w - cc.w) * inv_tt};
    const float zm = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    const float um = fmaxf(fmaxf(u[0], u[1]), fmaxf(u[2], u[3]));
    if (zm > st.ms) { st.ls *= __expf(st.ms - zm); st.ms = zm; }
    if (um > st.mt) { const float f = __expf(st.mt - um); st.lt *= f; st.dt *= f; st.mt = um; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      st.ls += __expf(z[e] - st.ms);
      const float w = __expf(u[e] - st.mt);
      st.lt += w;
      st.dt += w * z[e];
    }
  }
  // block reduction of the online states
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    OnlineState b;
    b.ms = __shfl_xor_sync(0xffffffffu, st.ms, o); b.ls = __shfl_xor_sync(0xffffffffu, st.ls, o);
    b.mt = __shfl_xor_sync(0xffffffffu, st.mt, o); b.lt = __shfl_xor_sync(0xffffffffu, st.lt, o);
    b.dt = __shfl_xor_sync(0xffffffffu, st.dt, o);
    online_merge(st, b);
  }
  __shared__ OnlineState sm[CE_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = st;
  __syncthreads();
  if (warp == 0) {
    st = (lane < CE_THREADS / 32) ? sm[lane] : OnlineState{-INFINITY, 0.f, -INFINITY, 0.f, 0.f};
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      OnlineState b;
      b.ms = __shfl_xor_sync(0xffffffffu, st.ms, o); b.ls = __shfl_xor_sync(0xffffffffu, st.ls, o);
      b.mt = __shfl_xor_sync(0xffffffffu, st.mt, o); b.lt = __shfl_xor_sync(0xffffffffu, st.lt, o);
      b.dt = __shfl_xor_sync(0xffffffffu, st.dt, o);
      online_merge(st, b);
    }
    if (lane == 0) {
      const float lse_s = st.ms + logf(st.ls);
      row_loss[i] = lse_s - st.dt / st.lt;
      stats[4 * i + 0] = lse_s;   // student log-sum-exp (of z/ts)
      stats[4 * i + 1] = st.mt;   // teacher max (of (t-c)/tt) for the partner row j
      stats[4 * i + 2] = st.lt;   // teacher sum exp
      stats[4 * i + 3] = 0.f;
    }
  }
}

// dz[i,k] = coef * (softmax(z_i/ts)_k - q_jk),  coef = gscale / (2R) / ts      (bf16: A operand of the head dgrad/wgrad)
__global__ void __launch_bounds__(CE_THREADS) dino_ce_bwd_kernel(const float* __restrict__ zs, const float* __restrict__ zt,
                                                                 const float* __restrict__ center,
                                                                 const float* __restrict__ stats,
                                                                 const float* __restrict__ gscale_ptr, float inv_ts,
                                                                 float inv_tt, bf16* __restrict__ dz, int R, int K) {
  const int i = blockIdx.x;
  const int j = (i + R) % (2 * R);
  const float lse_s = stats[4 * i + 0], mt = stats[4 * i + 1], inv_lt = 1.0f / stats[4 * i + 2];
  const float coef = (gscale_ptr ? *gscale_ptr : 1.0f) * inv_ts / (float)(2 * R);
  const float4* ps = reinterpret_cast<const float4*>(zs + (size_t)i * K);
  const float4* pt = reinterpret_cast<const float4*>(zt + (size_t)j * K);
  const float4* pc = reinterpret_cast<const float4*>(center);
  uint2* pd = reinterpret_cast<uint2*>(dz + (size_t)i * K);
  for (int c = threadIdx.x; c < (K >> 2); c += CE_THREADS) {
    const float4 a = ps[c];
    const float4 t = pt[c];
    const float4 cc = __ldg(pc + c);
    float d[4];
    d[0] = coef * (__expf(a.x * inv_ts - lse_s) - __expf((t.x - cc.x) * inv_tt - mt) * inv_lt);
    d[1] = coef * (__expf(a.y * inv_ts - lse_s) - __expf((t.y - cc.y) * inv_tt - mt) * inv_lt);
    d[2] = coef * (__expf(a.z * inv_ts - lse_s) - __expf((t.z - cc.z) * inv_tt - mt) * inv_lt);
    d[3] = coef * (__expf(a.w * inv_ts - lse_s) - __expf((t.w - cc.w) * inv_tt - mt) * inv_lt);
    pd[c] = make_uint2(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]));
  }
}

// sum of a float vector into out[0] scaled (deterministic single block)
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s * scale;
  }
}

// ---------------------------------------------------------------------------------------------------------
// segmentation loss: p = softmax(logits over 2 classes); loss = CE(p as logits, gt)  (double softmax)
// logits [N,2,H*W] f32, gt [N,H*W] f32 in {0,1}.  Forward writes per-block partial sums; backward is analytic.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void seg_pixel(float a, float b, int gt, float& loss, float& da) {
  const float m = fmaxf(a, b);
  const float ea = __expf(a - m), eb = __expf(b - m);
  const float p0 = ea / (ea + eb), p1 = eb / (ea + eb);
  const float mm = fmaxf(p0, p1);
  const float f0 = __expf(p0 - mm), f1 = __expf(p1 - mm);
  const float lse = mm + logf(f0 + f1);
  loss = lse - (gt ? p1 : p0);
  const float s0 = f0 / (f0 + f1), s1 = f1 / (f0 + f1);
  const float g0 = s0 - (gt ? 0.f : 1.f), g1 = s1 - (gt ? 1.f : 0.f);   // dL/dp0, dL/dp1
  da = (g0 - g1) * p0 * p1;                                               // dL/da ; dL/db = -da
}

__global__ void __launch_bounds__(256) seg_ce_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ gt,
                                                         float* __restrict__ partial, int n_img, int hw) {
  const size_t total = (size_t)n_img * hw;
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i % hw;
    float l, da;
    seg_pixel(logits[(img * 2) * hw + px], logits[(img * 2 + 1) * hw + px], gt[i] != 0.f, l, da);
    s += l;
  }
  __shared__ float sm[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) seg_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ gt,
                                                         const float* __restrict__ gscale_ptr, float* __restrict__ dlogits,
                                                         int n_img, int hw) {
  const size_t total = (size_t)n_img * hw;
  const float coef = (gscale_ptr ? *gscale_ptr : 1.0f) / (float)total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i % hw;
    float l, da;
    seg_pixel(logits[(img * 2) * hw + px], logits[(img * 2 + 1) * hw + px], gt[i] != 0.f, l, da);
    dlogits[(img * 2) * hw + px] = coef * da;
    dlogits[(img * 2 + 1) * hw + px] = -coef * da;
  }
}

}  // namespace ccd

using namespace ccd;

extern "C" int ccd_dino_ce_fwd(const float* zs, const float* zt, const float* center, float student_temp, float teacher_temp,
                               float* row_loss, float* stats, float* loss_out, int R, int K, void* stream) {
  if (!zs || !zt || !center || !row_loss || !stats || !loss_out || R <= 0 || K <= 0 || (K & 3)) return CCD_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  dino_ce_fwd_kernel<<<2 * R, CE_THREADS, 0, s>>>(zs, zt, center, 1.0f / student_temp, 1.0f / teacher_temp, row_loss, stats, R, K);
  CCD_LAUNCH_CHECK();
  reduce_sum_kernel<<<1, 1024, 0, s>>>(row_loss, 2 * R, 1.0f / (float)(2 * R), loss_out);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_dino_ce_bwd(const float* zs, const float* zt, const float* center, const float* stats,
                               const float* gscale_dev, float student_temp, float teacher_temp, void* dz_bf16, int R, int K,
                               void* stream) {
  if (!zs || !zt || !center || !stats || !dz_bf16 || R <= 0 || K <= 0 || (K & 3)) return CCD_ERR_ARG;
  dino_ce_bwd_kernel<<<2 * R, CE_THREADS, 0, (cudaStream_t)stream>>>(zs, zt, center, stats, gscale_dev, 1.0f / student_temp,
                                                                     1.0f / teacher_temp, (bf16*)dz_bf16, R, K);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_seg_ce_fwd(const float* logits, const float* gt, float* partial_ws, float* loss_out, int n_img, int hw,
                              void* stream) {
  if (!logits || !gt || !partial_ws || !loss_out || n_img <= 0 || hw <= 0) return CCD_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = 296;  // partial_ws must hold >= 296 floats
  seg_ce_fwd_kernel<<<blocks, 256, 0, s>>>(logits, gt, partial_ws, n_img, hw);
  CCD_LAUNCH_CHECK();
  reduce_sum_kernel<<<1, 1024, 0, s>>>(partial_ws, blocks, 1.0f / ((float)n_img * (float)hw), loss_out);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_seg_ce_bwd(const float* logits, const float* gt, const float* gscale_dev, float* dlogits, int n_img, int hw,
                              void* stream) {
  if (!logits || !gt || !dlogits || n_img <= 0 || hw <= 0) return CCD_ERR_ARG;
  seg_ce_bwd_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(logits, gt, gscale_dev, dlogits, n_img, hw);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
