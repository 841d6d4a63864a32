// Row-wise HBM-bound kernels of the CCD encoder / DINO head: LayerNorm fwd/bwd (eps 1e-6,
// Dino/modules/vision_transformer.py:108-110,247,250), bias-gradient column sums, L2 row normalisation and the
// weight-norm reparametrisation of DINOHead.last_layer (vision_transformer.py:313-316,326), multi-tensor cast / EMA
// (train.py:264-272).  One warp per row, float4 / 16-byte accesses, warp-shuffle reductions.
#include "ccd_common.cuh"

namespace ccd {

constexpr int LN_MAX_V4 = 4;  // E <= 512 -> at most 4 float4 per lane

// ---------------------------------------------------------------------------------------------------------
// LayerNorm forward: x f32 [rows,E] -> y (bf16 and/or f32)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y_bf16,
                                                            float* __restrict__ y_f32, int rows, int E, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nv = E >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * E);
  float4 v[LN_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      v[i] = xr[c];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (y_f32) reinterpret_cast<float4*>(y_f32 + (size_t)warp * E)[c] = o;
      if (y_bf16) {
        uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        reinterpret_cast<uint2*>(y_bf16 + (size_t)warp * E)[c] = pk;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward (recomputes mean/rstd from x): dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) + resid,
// g = dy*gamma.  HBM-bound (16.. bytes per element in+out): each warp owns rows r, r+W, ...; the loads of the NEXT
// row are issued before the reductions of the current one (software pipelining keeps ~2 rows of loads in flight
// per warp), dgamma/dbeta/dbias partial sums live in a per-warp shared-memory slab (plain read-modify-write, no
// atomics), reduced across the 8 warps at the end and added to global memory with one atomic per column per CTA.
// NV = float4 per lane (E/128 rounded up): 2 (E=192), 3 (E=384), 4 (E=512).
// ---------------------------------------------------------------------------------------------------------
template <bool DY_BF16, int NV>
struct LnRow {
  float4 x[NV];
  uint4 dy[NV];     // DY_BF16: .x/.y hold 4 bf16; else 4 floats
  float4 rs[NV];
};

template <bool DY_BF16, int NV>
__device__ __forceinline__ void ln_load_row(LnRow<DY_BF16, NV>& r, const float* __restrict__ x, const void* __restrict__ dy_,
                                            const float* __restrict__ resid, int row, int E, int nv, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      r.x[i] = reinterpret_cast<const float4*>(x + (size_t)row * E)[c];
      if constexpr (DY_BF16) {
        const uint2 pk = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_) + (size_t)row * E)[c];
        r.dy[i] = make_uint4(pk.x, pk.y, 0u, 0u);
      } else {
        r.dy[i] = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * E)[c];
      }
      r.rs[i] = resid ? reinterpret_cast<const float4*>(resid + (size_t)row * E)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <bool DY_BF16, int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const void* __restrict__ dy_, const float* __restrict__ resid,
                                                               float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               const float* __restrict__ bf16_seq_scale,
                                                               float* __restrict__ dbias_next, int rows, int E, float eps) {
  extern __shared__ float ln_acc[];                 // [8 warps][3][E]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warps_total = gridDim.x * 8;
  const int nv = E >> 2;
  float* my = ln_acc + (size_t)wib * 3 * E;
  pdl_launch_dependents();
  for (int i = lane; i < 3 * E; i += 32) my[i] = 0.f;
  __syncwarp();
  pdl_wait();
  float4 gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    gm[i] = (c < nv) ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_e = 1.0f / (float)E;

  int row = blockIdx.x * 8 + wib;
  LnRow<DY_BF16, NV> cur, nxt;
  if (row < rows) ln_load_row<DY_BF16, NV>(cur, x, dy_, resid, row, E, nv, lane);
  for (; row < rows; row += warps_total) {
    const int next_row = row + warps_total;
    if (next_row < rows) ln_load_row<DY_BF16, NV>(nxt, x, dy_, resid, next_row, E, nv, lane);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + i * 32 < nv) s += cur.x[i].x + cur.x[i].y + cur.x[i].z + cur.x[i].w;
    const float mean = warp_sum(s) * inv_e;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + i * 32 < nv) {
        cur.x[i].x -= mean; cur.x[i].y -= mean; cur.x[i].z -= mean; cur.x[i].w -= mean;
        q += cur.x[i].x * cur.x[i].x + cur.x[i].y * cur.x[i].y + cur.x[i].z * cur.x[i].z + cur.x[i].w * cur.x[i].w;
      }
    const float rstd = rsqrtf(warp_sum(q) * inv_e + eps);
    float m1 = 0.f, m2 = 0.f;
    float4 g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 d;
        if constexpr (DY_BF16) d = make_float4(bf16lo(cur.dy[i].x), bf16hi(cur.dy[i].x), bf16lo(cur.dy[i].y), bf16hi(cur.dy[i].y));
        else d = make_float4(__uint_as_float(cur.dy[i].x), __uint_as_float(cur.dy[i].y), __uint_as_float(cur.dy[i].z), __uint_as_float(cur.dy[i].w));
        cur.x[i].x *= rstd; cur.x[i].y *= rstd; cur.x[i].z *= rstd; cur.x[i].w *= rstd;   // xhat
        float4* ag = reinterpret_cast<float4*>(my) + c;                 // dgamma partial
        float4* ab = reinterpret_cast<float4*>(my + E) + c;             // dbeta partial
        float4 t = *ag;
        t.x += d.x * cur.x[i].x; t.y += d.y * cur.x[i].y; t.z += d.z * cur.x[i].z; t.w += d.w * cur.x[i].w;
        *ag = t;
        t = *ab;
        t.x += d.x; t.y += d.y; t.z += d.z; t.w += d.w;
        *ab = t;
        g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
        m1 += g[i].x + g[i].y + g[i].z + g[i].w;
        m2 += g[i].x * cur.x[i].x + g[i].y * cur.x[i].y + g[i].z * cur.x[i].z + g[i].w * cur.x[i].w;
      }
    }
    m1 = warp_sum(m1) * inv_e;
    m2 = warp_sum(m2) * inv_e;
    const float sc = bf16_seq_scale ? __ldg(bf16_seq_scale + (row >> 8)) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 o;
        o.x = rstd * (g[i].x - m1 - cur.x[i].x * m2) + cur.rs[i].x;
        o.y = rstd * (g[i].y - m1 - cur.x[i].y * m2) + cur.rs[i].y;
        o.z = rstd * (g[i].z - m1 - cur.x[i].z * m2) + cur.rs[i].z;
        o.w = rstd * (g[i].w - m1 - cur.x[i].w * m2) + cur.rs[i].w;
        if (dx_f32) reinterpret_cast<float4*>(dx_f32 + (size_t)row * E)[c] = o;
        if (dx_bf16) {
          // the bf16 copy feeds the NEXT residual branch's grads; DropPath scales that branch per sequence
          const uint2 pk = make_uint2(pack_bf16x2(o.x * sc, o.y * sc), pack_bf16x2(o.z * sc, o.w * sc));
          reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * E)[c] = pk;
          if (dbias_next) {   // column sums of the bf16 copy = bias gradient of the linear layer that consumes it as dY
            float4* an = reinterpret_cast<float4*>(my + 2 * E) + c;
            float4 t = *an;
            t.x += bf16lo(pk.x); t.y += bf16hi(pk.x); t.z += bf16lo(pk.y); t.w += bf16hi(pk.y);
            *an = t;
          }
        }
      }
    }
    cur = nxt;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ln_acc[(size_t)w * 3 * E + i];
    if (i < E) atomicAdd(dgamma + i, t);
    else if (i < 2 * E) atomicAdd(dbeta + i - E, t);
    else if (dbias_next) atomicAdd(dbias_next + i - 2 * E, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// out[c] += sum_r x[r,c]   (bias gradients), x bf16 [rows, cols], cols % 8 == 0, out pre-zeroed
// ---------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS_PER_BLOCK = 512;
constexpr int CS_BATCH = 8;           // independent 16 B loads in flight per thread
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, float* __restrict__ out, int rows,
                                                          int cols) {
  __shared__ float red[8][32][8];
  const int cg = blockIdx.x * 32 + threadIdx.x;  // column group of 8
  const int r0 = blockIdx.y * CS_ROWS_PER_BLOCK;
  const int r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cg * 8 < cols) {
    const bf16* base = x + (size_t)cg * 8;
    for (int r = r0 + threadIdx.y; r < r1; r += 8 * CS_BATCH) {
      uint4 v[CS_BATCH];
#pragma unroll
      for (int j = 0; j < CS_BATCH; ++j) {
        const int rr = r + 8 * j;
        v[j] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)rr * cols)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < CS_BATCH; ++j) {
        a[0] += bf16lo(v[j].x); a[1] += bf16hi(v[j].x); a[2] += bf16lo(v[j].y); a[3] += bf16hi(v[j].y);
        a[4] += bf16lo(v[j].z); a[5] += bf16hi(v[j].z); a[6] += bf16lo(v[j].w); a[7] += bf16hi(v[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = a[j];
  __syncthreads();
  // 256 threads reduce the 8 row-lanes of the 256 columns of this block: one atomic per column per block
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  if (col < cols) {
    float sum = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) sum += red[y][t >> 3][t & 7];
    atomicAdd(out + col, sum);
  }
}

// out[c] = sum_r x[r,c]   f32 [rows, cols] (teacher-centre batch sum, Dino/loss/Dino_loss.py:138); out pre-zeroed
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int rows,
                                                         int cols, int rows_per_block) {
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column
  if (c4 * 4 >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float4 a = make_float4(0, 0, 0, 0);
  for (int r = r0; r < r1; ++r) {
    const float4 v = reinterpret_cast<const float4*>(x + (size_t)r * cols)[c4];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  atomicAdd(out + 4 * c4 + 0, a.x); atomicAdd(out + 4 * c4 + 1, a.y);
  atomicAdd(out + 4 * c4 + 2, a.z); atomicAdd(out + 4 * c4 + 3, a.w);
}

// ---------------------------------------------------------------------------------------------------------
// L2 row normalisation (F.normalize, eps 1e-12) fwd/bwd over [rows, 256]; weight-norm of last_layer
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                         float* __restrict__ inv_norm, int rows, int cols) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float v = x[(size_t)row * cols + c]; s += v * v; }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < cols; c += 32) y[(size_t)row * cols + c] = __float2bfloat16(x[(size_t)row * cols + c] * inv);
  if (lane == 0) inv_norm[row] = inv;
}
// dx = inv * (dy - yhat * (yhat . dy)), yhat = x*inv (recomputed in fp32)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ inv_norm,
                                                         const float* __restrict__ dy, bf16* __restrict__ dx, int rows,
                                                         int cols) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += x[(size_t)row * cols + c] * inv * dy[(size_t)row * cols + c];
  dot = warp_sum(dot);
  for (int c = lane; c < cols; c += 32) {
    const float yh = x[(size_t)row * cols + c] * inv;
    dx[(size_t)row * cols + c] = __float2bfloat16(inv * (dy[(size_t)row * cols + c] - yh * dot));
  }
}
// w = v * g / ||v||  (row-wise over [K, cols]); writes bf16 w and inv_norm
__global__ void __launch_bounds__(256) weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                             bf16* __restrict__ w, float* __restrict__ inv_norm, int rows,
                                                             int cols) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { const float t = v[(size_t)row * cols + c]; s += t * t; }
  const float inv = rsqrtf(warp_sum(s));
  const float sc = g[row] * inv;
  for (int c = lane; c < cols; c += 32) w[(size_t)row * cols + c] = __float2bfloat16(v[(size_t)row * cols + c] * sc);
  if (lane == 0) inv_norm[row] = inv;
}
// dg = (dW . v) * inv ;  dv = g*inv * (dW - vhat * (dW . vhat)),  vhat = v*inv
__global__ void __launch_bounds__(256) weightnorm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                             const float* __restrict__ g, const float* __restrict__ inv_norm,
                                                             float* __restrict__ dv, float* __restrict__ dg, int rows, int cols) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot += dw[(size_t)row * cols + c] * v[(size_t)row * cols + c];
  dot = warp_sum(dot) * inv;  // dW . vhat
  const float sc = g[row] * inv;
  for (int c = lane; c < cols; c += 32)
    dv[(size_t)row * cols + c] = sc * (dw[(size_t)row * cols + c] - v[(size_t)row * cols + c] * inv * dot);
  if (lane == 0 && dg) dg[row] = dot;
}

// ---------------------------------------------------------------------------------------------------------
// multi-tensor kernels over a chunk table: int64 [n_chunks, 3] = (src_ptr, dst_ptr, n_elems <= 65536*?)
// ---------------------------------------------------------------------------------------------------------
enum MultiOp { MT_CAST_BF16 = 0, MT_EMA = 1, MT_SCALE = 2, MT_SQNORM = 3, MT_CLIP = 4 };
template <int OP>
__global__ void __launch_bounds__(256) multi_tensor_kernel(const long long* __restrict__ table, float a, float b) {
  const long long* e = table + (size_t)blockIdx.x * 3;
  const float* src = reinterpret_cast<const float*>(e[0]);
  const int n = (int)e[2];
  if constexpr (OP == MT_CAST_BF16) {
    bf16* dst = reinterpret_cast<bf16*>(e[1]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16(src[i]);
  } else if constexpr (OP == MT_EMA) {   // dst = a*dst + b*src   (teacher EMA, train.py:268-272)
    float* dst = reinterpret_cast<float*>(e[1]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = a * dst[i] + b * src[i];
  } else if constexpr (OP == MT_SCALE) {  // dst = a*src
    float* dst = reinterpret_cast<float*>(e[1]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = a * src[i];
  } else if constexpr (OP == MT_SQNORM) { // *dst += sum(src^2)   (dst = &sqnorm[tensor]; zero-filled by the caller)
    float* dst = reinterpret_cast<float*>(e[1]);
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += src[i] * src[i];
    __shared__ float sm[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sm[w];
      atomicAdd(dst, t);
    }
  } else {                                // per-parameter clip (Dino/modules/utils.py:132-141): src = &sqnorm[tensor],
    float* dst = reinterpret_cast<float*>(e[1]);   // dst = grad chunk; g *= clip/(||g||+1e-6) if that is < 1
    const float coef = a / (sqrtf(src[0]) + 1e-6f);
    if (coef < 1.0f)
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] *= coef;
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (size_t j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// centre EMA with the reference's divisor quirk: c = c*m + (sum / denom)*(1-m)   (Dino_loss.py:140-143)
__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ sum, float inv_denom, float m, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) center[i] = center[i] * m + sum[i] * inv_denom * (1.0f - m);
}

// im2col for the 4x4/stride-4 patch embedding (vision_transformer.py:126-131): x f32 [N,3,32,128] ->
// cols bf16 [N*256, 64] with k = c*16 + ky*4 + kx (conv weight [E,3,4,4] flattened), k in [48,64) zero padded.
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ cols, int n_img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (token, c, ky) -> 4 contiguous kx
  const int total = n_img * 256 * 16;
  if (idx >= total) return;
  const int part = idx & 15;           // 0..11 real (c*4+ky), 12..15 padding
  const int tok = idx >> 4;
  bf16* dst = cols + (size_t)tok * 64 + part * 4;
  if (part >= 12) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(0u, 0u);
    return;
  }
  const int img = tok >> 8, t = tok & 255, py = t >> 5, px = t & 31;
  const int c = part >> 2, ky = part & 3;
  const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)img * 3 + c) * 32 + (py * 4 + ky)) * 128 + px * 4);
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

}  // namespace ccd

using namespace ccd;

extern "C" int ccd_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                 int rows, int E, float eps, void* stream) {
  if (!x || !gamma || !beta || rows <= 0 || E <= 0 || (E & 3) || E > 512) return CCD_ERR_ARG;
  const int blocks = (rows + 7) / 8;
  CCD_CUDA_CHECK(launch_pdl(layernorm_fwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, (bf16*)y_bf16, y_f32,
                            rows, E, eps));
  return CCD_OK;
}

template <bool DY_BF16, int NV>
static int launch_ln_bwd(const float* x, const float* gamma, const void* dy, const float* resid, float* dx_f32, void* dx_bf16,
                         float* dgamma, float* dbeta, const float* bf16_seq_scale, float* dbias_next, int rows, int E, float eps,
                         cudaStream_t stream) {
  static bool attr_set = false;
  const int smem = 8 * 3 * E * (int)sizeof(float);
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(layernorm_bwd_kernel<DY_BF16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 512 * 4));
    attr_set = true;
  }
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;          // two resident CTAs per SM, each warp strides over its rows
  CCD_CUDA_CHECK(launch_pdl(layernorm_bwd_kernel<DY_BF16, NV>, dim3(blocks), dim3(256), (size_t)smem, stream, x, gamma, dy, resid, dx_f32,
                            (bf16*)dx_bf16, dgamma, dbeta, bf16_seq_scale, dbias_next, rows, E, eps));
  return CCD_OK;
}

extern "C" int ccd_layernorm_bwd(const float* x, const float* gamma, const void* dy, int dy_is_bf16, const float* resid,
                                 float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, const float* bf16_seq_scale,
                                 float* dbias_next, int rows, int E, float eps, void* stream) {
  if (!x || !gamma || !dy || !dgamma || !dbeta || rows <= 0 || (E & 3) || E > 512) return CCD_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int nvl = (E / 4 + 31) / 32;
#define CCD_LN_BWD(B, N) return launch_ln_bwd<B, N>(x, gamma, dy, resid, dx_f32, dx_bf16, dgamma, dbeta, bf16_seq_scale, dbias_next, rows, E, eps, s)
  if (dy_is_bf16) {
    if (nvl <= 2) CCD_LN_BWD(true, 2);
    if (nvl == 3) CCD_LN_BWD(true, 3);
    CCD_LN_BWD(true, 4);
  } else {
    if (nvl <= 2) CCD_LN_BWD(false, 2);
    if (nvl == 3) CCD_LN_BWD(false, 3);
    CCD_LN_BWD(false, 4);
  }
#undef CCD_LN_BWD
}

extern "C" int ccd_colsum_bf16(const void* x, float* out, int rows, int cols, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || (cols & 7)) return CCD_ERR_ARG;
  dim3 grid((cols + 255) / 256, (rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK);
  colsum_bf16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const bf16*)x, out, rows, cols);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

// out[c] += sum_r v[r] * W[r, c]: one thread per column, rows split over blockIdx.y (coalesced over c; a few hundred KB at most)
__global__ void __launch_bounds__(128) vecmat_add_f32_kernel(const float* __restrict__ v, const float* __restrict__ W,
                                                             float* __restrict__ out, int rows, int cols, int rpb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc = fmaf(v[r], W[(size_t)r * cols + c], acc);
  atomicAdd(out + c, acc);
}

// C[M,N] = op(A) B in fp32 for the SMALL constant operators of the path (the 256 x 256 bicubic pos-embed resample):
// op(A) = A [M,K] or A^T with A stored [K,M]; B [K,N]; one thread per output element, coalesced over n.
__global__ void __launch_bounds__(128) smallmm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                          int M, int N, int K, int trans_a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  if (!trans_a) {
    for (int k = 0; k < K; ++k) acc = fmaf(A[(size_t)m * K + k], B[(size_t)k * N + n], acc);
  } else {
    for (int k = 0; k < K; ++k) acc = fmaf(A[(size_t)k * M + m], B[(size_t)k * N + n], acc);
  }
  C[(size_t)m * N + n] = acc;
}

extern "C" int ccd_smallmm_f32(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || M > 65535) return CCD_ERR_ARG;
  smallmm_f32_kernel<<<dim3((N + 127) / 128, M), 128, 0, (cudaStream_t)stream>>>(A, B, C, M, N, K, trans_a ? 1 : 0);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_vecmat_add_f32(const float* v, const float* W, float* out, int rows, int cols, void* stream) {
  if (!v || !W || !out || rows <= 0 || cols <= 0) return CCD_ERR_ARG;
  const int rpb = 32;
  dim3 grid((cols + 127) / 128, (rows + rpb - 1) / rpb);
  vecmat_add_f32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(v, W, out, rows, cols, rpb);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_colsum_f32(const float* x, float* out, int rows, int cols, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || (cols & 3)) return CCD_ERR_ARG;
  const int rpb = 64;
  dim3 grid((cols / 4 + 255) / 256, (rows + rpb - 1) / rpb);
  colsum_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, rows, cols, rpb);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_l2norm_fwd(const float* x, void* y_bf16, float* inv_norm, int rows, int cols, void* stream) {
  if (!x || !y_bf16 || !inv_norm || rows <= 0) return CCD_ERR_ARG;
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y_bf16, inv_norm, rows, cols);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_l2norm_bwd(const float* x, const float* inv_norm, const float* dy, void* dx_bf16, int rows, int cols,
                              void* stream) {
  if (!x || !dy || !dx_bf16 || !inv_norm || rows <= 0) return CCD_ERR_ARG;
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, inv_norm, dy, (bf16*)dx_bf16, rows, cols);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_weightnorm_fwd(const float* v, const float* g, void* w_bf16, float* inv_norm, int rows, int cols,
                                  void* stream) {
  if (!v || !g || !w_bf16 || !inv_norm || rows <= 0) return CCD_ERR_ARG;
  weightnorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(v, g, (bf16*)w_bf16, inv_norm, rows, cols);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv,
                                  float* dg, int rows, int cols, void* stream) {
  if (!dw || !v || !g || !inv_norm || !dv || rows <= 0) return CCD_ERR_ARG;
  weightnorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dw, v, g, inv_norm, dv, dg, rows, cols);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_multi_tensor(int op, const void* table_dev, int n_chunks, float a, float b, void* stream) {
  if (!table_dev || n_chunks <= 0) return CCD_ERR_ARG;
  const long long* t = (const long long*)table_dev;
  cudaStream_t s = (cudaStream_t)stream;
  switch (op) {
    case MT_CAST_BF16: multi_tensor_kernel<MT_CAST_BF16><<<n_chunks, 256, 0, s>>>(t, a, b); break;
    case MT_EMA: multi_tensor_kernel<MT_EMA><<<n_chunks, 256, 0, s>>>(t, a, b); break;
    case MT_SCALE: multi_tensor_kernel<MT_SCALE><<<n_chunks, 256, 0, s>>>(t, a, b); break;
    case MT_SQNORM: multi_tensor_kernel<MT_SQNORM><<<n_chunks, 256, 0, s>>>(t, a, b); break;
    case MT_CLIP: multi_tensor_kernel<MT_CLIP><<<n_chunks, 256, 0, s>>>(t, a, b); break;
    default: return CCD_ERR_ARG;
  }
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (!src || !dst || n <= 0) return CCD_ERR_ARG;
  const long long threads = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, (size_t)n);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_center_ema(float* center, const float* sum, float denom, float momentum, int n, void* stream) {
  if (!center || !sum || n <= 0 || denom <= 0.f) return CCD_ERR_ARG;
  center_ema_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(center, sum, 1.0f / denom, momentum, n);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_patch_im2col(const float* x, void* cols_bf16, int n_img, void* stream) {
  if (!x || !cols_bf16 || n_img <= 0) return CCD_ERR_ARG;
  const int total = n_img * 256 * 16;
  patch_im2col_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)cols_bf16, n_img);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
