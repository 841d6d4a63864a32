// Character-segment path of CCD (integer / index work, bit-exact against the oracle):
//   * connected-component labelling of the 32x128 text mask        (label_cluster, Dino/utils/DBSCAN.py:61-103;
//                                                                    driver loop Dino/model/dino_vision.py:59-71)
//   * affine warp of the cluster maps / GT mask to the second view  (dino_vision.py:72-78, train.py:234-236)
//   * mask-guided character pooling + ragged row selection          (ABIDINOModel.attention, dino_vision.py:38-49,80-87)
// The reference does the labelling on the CPU with a float64 [B,26,32,128] host round trip; here one CTA labels one
// image in shared memory and the cluster maps live in HBM as ONE uint32 bitmask per pixel (bit s = slot s), for both
// views (the warped view is not a partition: neighbouring characters can both exceed the 0.1 threshold at a pixel).
#include "ccd_common.cuh"

namespace ccd {

constexpr int IMG_H = 32, IMG_W = 128, IMG_PX = IMG_H * IMG_W;
constexpr int SLOTS = 26;
constexpr int MIN_AREA = 30;          // DBSCAN.py:88
constexpr int TOK_H = 8, TOK_W = 32;  // token grid (vision_transformer.py:237-238)

// ---------------------------------------------------------------------------------------------------------
// connected components: union-find in shared memory, 8-connectivity, roots = smallest raster index
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(volatile int* lab, int x) {
  int p = lab[x];
  while (p != x) { x = p; p = lab[x]; }
  return x;
}
__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
  while (true) {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b : hang the larger root under the smaller
    const int old = atomicMin(&lab[a], b);
    if (old == a) return;                         // a was still a root: linked
    a = old;                                      // somebody else linked a meanwhile: retry from there
  }
}

// mode 0: src = f32 mask [B,32,128] (foreground = non-zero)
// mode 1: src = f32 seg logits [*,2,32,128]; foreground = softmax[:,1] > 0.5 == logit1 > logit0 (dino_vision.py:64-66)
__global__ void __launch_bounds__(256) ccl_label_kernel(const float* __restrict__ src, int mode, unsigned* __restrict__ bits,
                                                        unsigned char* __restrict__ compact, int* __restrict__ n_comp) {
  extern __shared__ int ccl_smem[];
  int* lab = ccl_smem;               // [4096]
  int* area = lab + IMG_PX;          // [4096]
  int* sumx = area + IMG_PX;         // [4096]  (later: slot map)
  __shared__ int warp_tot[8];
  __shared__ int kroot[SLOTS], karea[SLOTS], ksumx[SLOTS], kslot[SLOTS];
  __shared__ int n_kept_s;
  const int b = blockIdx.x, tid = threadIdx.x;

  for (int p = tid; p < IMG_PX; p += 256) {
    bool fg;
    if (mode == 0) fg = src[(size_t)b * IMG_PX + p] != 0.f;
    else fg = src[((size_t)b * 2 + 1) * IMG_PX + p] > src[((size_t)b * 2) * IMG_PX + p];
    lab[p] = fg ? p : -1;
    area[p] = 0;
    sumx[p] = 0;
  }
  __syncthreads();
  for (int p = tid; p < IMG_PX; p += 256) {
    if (lab[p] < 0) continue;
    const int y = p >> 7, x = p & 127;
    if (x > 0 && lab[p - 1] >= 0) uf_union(lab, p, p - 1);
    if (y > 0) {
      if (lab[p - IMG_W] >= 0) uf_union(lab, p, p - IMG_W);
      if (x > 0 && lab[p - IMG_W - 1] >= 0) uf_union(lab, p, p - IMG_W - 1);
      if (x < IMG_W - 1 && lab[p - IMG_W + 1] >= 0) uf_union(lab, p, p - IMG_W + 1);
    }
  }
  __syncthreads();
  for (int p = tid; p < IMG_PX; p += 256) {
    if (lab[p] < 0) continue;
    const int r = uf_find(lab, p);
    atomicAdd(&area[r], 1);
    atomicAdd(&sumx[r], p & 127);
  }
  __syncthreads();
  for (int p = tid; p < IMG_PX; p += 256)
    if (lab[p] >= 0) lab[p] = uf_find(lab, p);   // flatten (roots keep pointing at themselves)
  __syncthreads();

  // candidates in label (= root raster) order: thread t owns pixels [16t, 16t+16)
  int local = 0;
  for (int k = 0; k < 16; ++k) {
    const int p = tid * 16 + k;
    local += (lab[p] == p && area[p] >= MIN_AREA) ? 1 : 0;
  }
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
  int rank = base + incl - local;
  if (tid == 255) n_kept_s = min(base + incl, SLOTS);
  for (int k = 0; k < 16; ++k) {
    const int p = tid * 16 + k;
    if (lab[p] == p && area[p] >= MIN_AREA) {
      if (rank < SLOTS) { kroot[rank] = p; karea[rank] = area[p]; ksumx[rank] = sumx[p]; }   // first 26 (DBSCAN.py:92-93)
      ++rank;
    }
  }
  __syncthreads();
  const int n_kept = n_kept_s;
  // order by mean column = sumx/area (exact rational compare); ties keep label order (np.argsort, DBSCAN.py:94)
  if (tid < n_kept) {
    int pos = 0;
    const long long sk = ksumx[tid], ak = karea[tid];
    for (int l = 0; l < n_kept; ++l) {
      const long long lhs = (long long)ksumx[l] * ak, rhs = sk * (long long)karea[l];
      if (lhs < rhs || (lhs == rhs && l < tid)) ++pos;
    }
    kslot[tid] = pos;
  }
  for (int p = tid; p < IMG_PX; p += 256) sumx[p] = 0;   // reuse as root -> slot+1 map
  __syncthreads();
  if (tid < n_kept) sumx[kroot[tid]] = kslot[tid] + 1;
  __syncthreads();
  for (int p = tid; p < IMG_PX; p += 256) {
    const int s = (lab[p] >= 0) ? sumx[lab[p]] : 0;
    if (bits) bits[(size_t)b * IMG_PX + p] = s ? (1u << (s - 1)) : 0u;
    if (compact) compact[(size_t)b * IMG_PX + p] = (unsigned char)s;
  }
  if (tid == 0 && n_comp) n_comp[b] = n_kept;
}

// ---------------------------------------------------------------------------------------------------------
// affine_grid + bilinear grid_sample (zeros padding, align_corners=False) + "> 0.1"
// ---------------------------------------------------------------------------------------------------------
struct WarpTaps { int x0, y0; float w[4]; };   // taps: (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1)

__device__ __forceinline__ float linspace_m1_1(int i, int steps) {   // ATen linspace(-1, 1, steps), fp32
  const float step = 2.0f / (float)(steps - 1);
  return (i < steps / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(steps - 1 - i));
}
__device__ __forceinline__ WarpTaps warp_taps(const float* __restrict__ th, int y, int x) {
  const float bx = linspace_m1_1(x, IMG_W) * ((float)(IMG_W - 1) / (float)IMG_W);
  const float by = linspace_m1_1(y, IMG_H) * ((float)(IMG_H - 1) / (float)IMG_H);
  const float gx = bx * th[0] + by * th[1] + th[2];
  const float gy = bx * th[3] + by * th[4] + th[5];
  const float ix = ((gx + 1.0f) * (float)IMG_W - 1.0f) * 0.5f;
  const float iy = ((gy + 1.0f) * (float)IMG_H - 1.0f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  WarpTaps t;
  t.x0 = (int)fx; t.y0 = (int)fy;
  const float ex = (fx + 1.0f) - ix, wx = ix - fx, ey = (fy + 1.0f) - iy, wy = iy - fy;
  t.w[0] = ex * ey; t.w[1] = wx * ey; t.w[2] = ex * wy; t.w[3] = wx * wy;
  return t;
}
__device__ __forceinline__ bool in_img(int y, int x) { return (unsigned)y < (unsigned)IMG_H && (unsigned)x < (unsigned)IMG_W; }

__global__ void __launch_bounds__(256) warp_bits_kernel(const unsigned* __restrict__ src, const float* __restrict__ theta,
                                                        unsigned* __restrict__ dst, int n_img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * IMG_PX) return;
  const int b = idx / IMG_PX, p = idx % IMG_PX;
  const WarpTaps t = warp_taps(theta + b * 9, p >> 7, p & 127);
  const unsigned* s = src + (size_t)b * IMG_PX;
  unsigned m[4];
  m[0] = in_img(t.y0, t.x0) ? s[t.y0 * IMG_W + t.x0] : 0u;
  m[1] = in_img(t.y0, t.x0 + 1) ? s[t.y0 * IMG_W + t.x0 + 1] : 0u;
  m[2] = in_img(t.y0 + 1, t.x0) ? s[(t.y0 + 1) * IMG_W + t.x0] : 0u;
  m[3] = in_img(t.y0 + 1, t.x0 + 1) ? s[(t.y0 + 1) * IMG_W + t.x0 + 1] : 0u;
  unsigned any = m[0] | m[1] | m[2] | m[3], out = 0u;
  while (any) {
    const int sl = __ffs(any) - 1;
    any &= any - 1;
    float v = 0.f;     // same tap order as ATen's grid_sampler: nw, ne, sw, se
    if ((m[0] >> sl) & 1u) v += t.w[0];
    if ((m[1] >> sl) & 1u) v += t.w[1];
    if ((m[2] >> sl) & 1u) v += t.w[2];
    if ((m[3] >> sl) & 1u) v += t.w[3];
    if (v > 0.1f) out |= 1u << sl;
  }
  dst[idx] = out;
}

__global__ void __launch_bounds__(256) warp_mask_kernel(const float* __restrict__ src, const float* __restrict__ theta,
                                                        float* __restrict__ dst, int n_img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * IMG_PX) return;
  const int b = idx / IMG_PX, p = idx % IMG_PX;
  const WarpTaps t = warp_taps(theta + b * 9, p >> 7, p & 127);
  const float* s = src + (size_t)b * IMG_PX;
  float v = 0.f;
  if (in_img(t.y0, t.x0)) v += s[t.y0 * IMG_W + t.x0] * t.w[0];
  if (in_img(t.y0, t.x0 + 1)) v += s[t.y0 * IMG_W + t.x0 + 1] * t.w[1];
  if (in_img(t.y0 + 1, t.x0)) v += s[(t.y0 + 1) * IMG_W + t.x0] * t.w[2];
  if (in_img(t.y0 + 1, t.x0 + 1)) v += s[(t.y0 + 1) * IMG_W + t.x0 + 1] * t.w[3];
  dst[idx] = v > 0.1f ? 1.0f : 0.0f;
}

// bits [N,32,128] <-> dense one-hot f32 [N,26,32,128]  (API compatibility: student_output['zero'])
__global__ void __launch_bounds__(256) bits_to_dense_kernel(const unsigned* __restrict__ bits, float* __restrict__ dense, int n_img) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_img * SLOTS * IMG_PX) return;
  const int p = idx % IMG_PX;
  const int s = (idx / IMG_PX) % SLOTS;
  const size_t b = idx / ((size_t)SLOTS * IMG_PX);
  dense[idx] = ((bits[b * IMG_PX + p] >> s) & 1u) ? 1.0f : 0.0f;
}
__global__ void __launch_bounds__(256) dense_to_bits_kernel(const float* __restrict__ dense, unsigned* __restrict__ bits, int n_img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * IMG_PX) return;
  const int b = idx / IMG_PX, p = idx % IMG_PX;
  unsigned m = 0u;
  for (int s = 0; s < SLOTS; ++s)
    if (dense[((size_t)b * SLOTS + s) * IMG_PX + p] != 0.f) m |= 1u << s;
  bits[idx] = m;
}

// ---------------------------------------------------------------------------------------------------------
// pooling weights.  bilinear /4 with align_corners=False == mean of the central 2x2 of each 4x4 cell, so the
// per-token weight of a slot is count(set among 4 pixels)/4; normalised by the slot total over the 256 tokens.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void token_masks(const unsigned* __restrict__ bimg, int t, unsigned (&m)[4]) {
  const int ty = t >> 5, tx = t & 31;
  const unsigned* r0 = bimg + (ty * 4 + 1) * IMG_W + tx * 4 + 1;
  m[0] = r0[0]; m[1] = r0[1]; m[2] = r0[IMG_W]; m[3] = r0[IMG_W + 1];
}

// per image: tot[slot] (in quarter units) and occupied-slot count; first-view images also give the row count
__global__ void __launch_bounds__(256) char_index_kernel(const unsigned* __restrict__ bits, int* __restrict__ tot4,
                                                         int* __restrict__ cnt, int n_view) {
  __shared__ int s_tot[SLOTS];
  const int n = blockIdx.x, t = threadIdx.x;
  if (t < SLOTS) s_tot[t] = 0;
  __syncthreads();
  unsigned m[4];
  token_masks(bits + (size_t)n * IMG_PX, t, m);
  unsigned any = m[0] | m[1] | m[2] | m[3];
  while (any) {
    const int sl = __ffs(any) - 1;
    any &= any - 1;
    const int c = ((m[0] >> sl) & 1) + ((m[1] >> sl) & 1) + ((m[2] >> sl) & 1) + ((m[3] >> sl) & 1);
    atomicAdd(&s_tot[sl], c);
  }
  __syncthreads();
  if (t < SLOTS) tot4[n * SLOTS + t] = s_tot[t];
  if (t == 0 && n < n_view) {
    int occ = 0;
    for (int s = 0; s < SLOTS; ++s) occ += s_tot[s] > 0;
    const int length = min(max(occ, 3), SLOTS);          // clamp(index.sum, 3, 26)   dino_vision.py:83
    cnt[n] = min(length, SLOTS - 1) + 1;                 // slots 0..length inclusive, at most 26
  }
}
// exclusive scan of cnt -> row offsets; total rows per view R -> offs[n_view]
__global__ void __launch_bounds__(1024) char_offsets_kernel(const int* __restrict__ cnt, int* __restrict__ offs, int n_view) {
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_view; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n_view) ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wb += wsum[w];
    const int c0 = carry;
    if (i < n_view) offs[i] = c0 + wb + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offs[n_view] = carry;
}

// rows[off(b)+slot] (view 1) / rows[R+off(b)+slot] (view 2) = sum_t w[slot][t] * tokens[n][t][:]
template <typename TokT>
__global__ void __launch_bounds__(256) char_pool_fwd_kernel(const TokT* __restrict__ tokens, const unsigned* __restrict__ bits,
                                                            const int* __restrict__ tot4, const int* __restrict__ cnt,
                                                            const int* __restrict__ offs, float* __restrict__ rows, int n_view,
                                                            int E) {
  extern __shared__ float pool_smem[];                 // acc [26][E]
  __shared__ unsigned s_mask[256][4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int b = (n < n_view) ? n : n - n_view;
  const int nrows = cnt[b];
  const int R = offs[n_view];
  const int row0 = offs[b] + ((n < n_view) ? 0 : R);
  for (int i = tid; i < SLOTS * E; i += 256) pool_smem[i] = 0.f;
  {
    unsigned m[4];
    token_masks(bits + (size_t)n * IMG_PX, tid, m);
    s_mask[tid][0] = m[0]; s_mask[tid][1] = m[1]; s_mask[tid][2] = m[2]; s_mask[tid][3] = m[3];
  }
  __syncthreads();
  const unsigned sel = (nrows >= 32) ? 0xffffffffu : ((1u << nrows) - 1u);   // only selected slots are needed
  for (int t = 0; t < 256; ++t) {
    const unsigned m0 = s_mask[t][0], m1 = s_mask[t][1], m2 = s_mask[t][2], m3 = s_mask[t][3];
    unsigned any = (m0 | m1 | m2 | m3) & sel;
    if (!any) continue;                                                      // block-uniform
    const TokT* tok = tokens + ((size_t)n * 256 + t) * E;
    while (any) {
      const int sl = __ffs(any) - 1;
      any &= any - 1;
      const float c = 0.25f * (float)(((m0 >> sl) & 1) + ((m1 >> sl) & 1) + ((m2 >> sl) & 1) + ((m3 >> sl) & 1));
      for (int e = tid; e < E; e += 256) pool_smem[sl * E + e] += c * (float)tok[e];
    }
  }
  __syncthreads();
  for (int i = tid; i < nrows * E; i += 256) {
    const int sl = i / E;
    const int tt = tot4[n * SLOTS + sl];
    rows[(size_t)(row0 + sl) * E + (i - sl * E)] = tt > 0 ? pool_smem[i] / (0.25f * (float)tt) : 0.f;   // 0/0 -> 0 (:45)
  }
}

// dtokens[n][t][:] = sum_{slot selected, in token t} w[slot][t] * drows[row(slot)][:]
__global__ void __launch_bounds__(256) char_pool_bwd_kernel(const float* __restrict__ drows, const unsigned* __restrict__ bits,
                                                            const int* __restrict__ tot4, const int* __restrict__ cnt,
                                                            const int* __restrict__ offs, float* __restrict__ dtokens, int n_view,
                                                            int E) {
  extern __shared__ float pool_smem[];                 // drows of this image, pre-divided by tot: [26][E]
  __shared__ unsigned s_mask[256][4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int b = (n < n_view) ? n : n - n_view;
  const int nrows = cnt[b];
  const int R = offs[n_view];
  const int row0 = offs[b] + ((n < n_view) ? 0 : R);
  for (int i = tid; i < nrows * E; i += 256) {
    const int sl = i / E;
    const int tt = tot4[n * SLOTS + sl];
    pool_smem[i] = tt > 0 ? drows[(size_t)(row0 + sl) * E + (i - sl * E)] / (0.25f * (float)tt) : 0.f;
  }
  {
    unsigned m[4];
    token_masks(bits + (size_t)n * IMG_PX, tid, m);
    s_mask[tid][0] = m[0]; s_mask[tid][1] = m[1]; s_mask[tid][2] = m[2]; s_mask[tid][3] = m[3];
  }
  __syncthreads();
  const unsigned sel = (nrows >= 32) ? 0xffffffffu : ((1u << nrows) - 1u);
  for (int t = 0; t < 256; ++t) {
    const unsigned m0 = s_mask[t][0], m1 = s_mask[t][1], m2 = s_mask[t][2], m3 = s_mask[t][3];
    const unsigned any0 = (m0 | m1 | m2 | m3) & sel;
    float* dt = dtokens + ((size_t)n * 256 + t) * E;
    for (int e = tid; e < E; e += 256) {
      float acc = 0.f;
      unsigned any = any0;
      while (any) {
        const int sl = __ffs(any) - 1;
        any &= any - 1;
        const float c = 0.25f * (float)(((m0 >> sl) & 1) + ((m1 >> sl) & 1) + ((m2 >> sl) & 1) + ((m3 >> sl) & 1));
        acc += c * pool_smem[sl * E + e];
      }
      dt[e] = acc;
    }
  }
}

// new_index[b][s] = s <= length(b)   (bool as uint8; dino_vision.py:84-85)
__global__ void new_index_kernel(const int* __restrict__ cnt, unsigned char* __restrict__ out, int n_view) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_view * SLOTS) out[i] = (i % SLOTS) < cnt[i / SLOTS] ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------------------
// Text-mask generation: 2-means over the grey levels of one crop + border-majority polarity rule
// (clusterpixels, mask_create/generate_mask.py:13-29 = Dino/utils/kmeans.py:8-24).
//
// The reference runs scipy.cluster.vq.kmeans(k = 2) -- Lloyd iterations from 20 random initialisations, best distortion
// kept -- and then assigns every pixel to its nearest centroid (vq).  In one dimension the partition Lloyd converges to is a
// threshold on the grey level, and the minimum-distortion one is found EXACTLY from the 256-bin histogram: maximise
// s0^2/n0 + s1^2/n1 over the 255 thresholds (prefix sums).  One CTA per image:
//   histogram (shared-memory atomics) -> block-wide prefix sums -> arg-max threshold -> centroids c0 < c1 -> code = grey > (c0+c1)/2
//   -> sums of the code over the first / last column and row -> flip when >= 3 of them exceed half (the rule that makes the
//   text the 1-cluster whichever centroid came first) -> mask f32 {0,1} [H,W], directly consumable by ccd_ccl_label mode 0.
// Centroid order is canonical (dark = 0, bright = 1); the reference's order depends on its random initialisation, which only
// shows when exactly two of the four border sums exceed half (its answer then depends on the seed).  Constant images give an
// all-zero mask (scipy drops the empty cluster).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kmeans_mask_kernel(const uint8_t* __restrict__ grey, float* __restrict__ mask, int H, int W) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long pre_n[256], pre_s[256];      // inclusive prefix counts / grey-level sums
  __shared__ double best_j[8];
  __shared__ int best_t[8];
  __shared__ unsigned int border[4];                          // code sums: first col, last col, first row, last row
  __shared__ float s_mid;
  __shared__ int s_flip, s_const;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int npx = H * W;
  const uint8_t* img = grey + (size_t)blockIdx.x * npx;
  float* out = mask + (size_t)blockIdx.x * npx;
  hist[tid] = 0u;
  if (tid < 4) border[tid] = 0u;
  __syncthreads();
  for (int i = tid; i < npx; i += 256) atomicAdd(&hist[img[i]], 1u);
  __syncthreads();
  // inclusive scan of (count, count * level) over the 256 bins: warp shuffles, then the 8 warp totals
  unsigned long long n = hist[tid], sm = (unsigned long long)hist[tid] * (unsigned long long)tid;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n2 = __shfl_up_sync(0xffffffffu, n, o), s2 = __shfl_up_sync(0xffffffffu, sm, o);
    if (lane >= o) { n += n2; sm += s2; }
  }
  pre_n[tid] = n;
  pre_s[tid] = sm;
  __syncthreads();
  unsigned long long off_n = 0, off_s = 0;
  for (int w = 0; w < wrp; ++w) { off_n += pre_n[w * 32 + 31]; off_s += pre_s[w * 32 + 31]; }
  __syncthreads();
  n += off_n;
  sm += off_s;
  pre_n[tid] = n;
  pre_s[tid] = sm;
  __syncthreads();
  const unsigned long long N = pre_n[255], S = pre_s[255];
  // threshold t = tid: class 0 = levels <= t.  J(t) = s0^2/n0 + s1^2/n1 in double (s <= 255 * 2^20: squares are exact)
  double j = -1.0;
  if (tid < 255 && n > 0 && n < N) {
    const double s0 = (double)sm, s1 = (double)(S - sm);
    j = s0 * s0 / (double)n + s1 * s1 / (double)(N - n);
  }
  int t = tid;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {                         // arg-max, ties -> smallest threshold
    const double j2 = __shfl_xor_sync(0xffffffffu, j, o);
    const int t2 = __shfl_xor_sync(0xffffffffu, t, o);
    if (j2 > j || (j2 == j && t2 < t)) { j = j2; t = t2; }
  }
  if (lane == 0) { best_j[wrp] = j; best_t[wrp] = t; }
  __syncthreads();
  if (tid == 0) {
    double bj = best_j[0];
    int bt = best_t[0];
    for (int w = 1; w < 8; ++w)
      if (best_j[w] > bj || (best_j[w] == bj && best_t[w] < bt)) { bj = best_j[w]; bt = best_t[w]; }
    if (bj < 0.0) {                                           // a single grey level: one cluster, every code 0
      s_const = 1;
      s_mid = 256.0f;
    } else {
      const double c0 = (double)pre_s[bt] / (double)pre_n[bt];
      const double c1 = (double)(S - pre_s[bt]) / (double)(N - pre_n[bt]);
      s_const = 0;
      s_mid = (float)(0.5 * (c0 + c1));                       // |v - c1| < |v - c0|  <=>  v > (c0 + c1) / 2  (ties -> code 0, as vq's arg-min)
    }
  }
  __syncthreads();
  const float mid = s_mid;
  // border sums of the code (generate_mask.py:21-24)
  unsigned int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  for (int y = tid; y < H; y += 256) {
    b0 += (float)img[y * W] > mid;
    b1 += (float)img[y * W + W - 1] > mid;
  }
  for (int x = tid; x < W; x += 256) {
    b2 += (float)img[x] > mid;
    b3 += (float)img[(H - 1) * W + x] > mid;
  }
  if (b0) atomicAdd(&border[0], b0);
  if (b1) atomicAdd(&border[1], b1);
  if (b2) atomicAdd(&border[2], b2);
  if (b3) atomicAdd(&border[3], b3);
  __syncthreads();
  if (tid == 0) {
    const int num = (int)(border[2] > (unsigned)(W / 2)) + (int)(border[3] > (unsigned)(W / 2)) + (int)(border[0] > (unsigned)(H / 2)) +
                    (int)(border[1] > (unsigned)(H / 2));
    s_flip = (num >= 3) ? 1 : 0;
  }
  __syncthreads();
  const bool flip = s_flip != 0;
  for (int i = tid; i < npx; i += 256) {
    const bool code = (float)img[i] > mid;
    out[i] = (code != flip) ? 1.0f : 0.0f;
  }
}


// ---------------------------------------------------------------------------------------------------------
// Affine theta of the irregular view (datasetsupervised_kmeans.py:60-71; SURVEY section 8f #4, second slice -- the algebra only).
// The dataset draws an imgaug Affine, takes the INVERSE pixel-space matrix M_inv of the warp it applied to the source-size
// image (`matric[0]._inv_matrix`), rescales it to the 128 x 32 network input and re-expresses it in grid_sample's normalised
// coordinates:   metric = W^-1 M_inv W,   theta = N metric N^-1,   W = diag(src_w / img_w, src_h / img_h, 1),
//                N = [[2/(img_w-1), 0, -1], [0, 2/(img_h-1), -1], [0, 0, 1]]          (float64, stored as float32).
// One thread per sample.  Drawing the augmentation parameters and building M_inv stay with imgaug (third party, not in this
// image): this kernel replaces lines 63-71 for a batch of matrices.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double (&a)[9], const double (&b)[9], double (&c)[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__global__ void affine_theta_kernel(const double* __restrict__ m_inv, const int* __restrict__ src_hw, float* __restrict__ theta, int n,
                                    int img_h, int img_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ws = (double)src_hw[2 * i + 1] / (double)img_w, hs = (double)src_hw[2 * i] / (double)img_h;
  const double Wi[9] = {1.0 / ws, 0, 0, 0, 1.0 / hs, 0, 0, 0, 1}, W[9] = {ws, 0, 0, 0, hs, 0, 0, 0, 1};
  const double N[9] = {2.0 / (img_w - 1), 0, -1, 0, 2.0 / (img_h - 1), -1, 0, 0, 1};
  const double Ni[9] = {(img_w - 1) / 2.0, 0, (img_w - 1) / 2.0, 0, (img_h - 1) / 2.0, (img_h - 1) / 2.0, 0, 0, 1};
  double M[9], t0[9], metric[9], t1[9], th[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = m_inv[9 * i + k];
  mat3_mul(Wi, M, t0);
  mat3_mul(t0, W, metric);
  mat3_mul(N, metric, t1);
  mat3_mul(t1, Ni, th);
#pragma unroll
  for (int k = 0; k < 9; ++k) theta[9 * i + k] = (float)th[k];
}

}  // namespace ccd

using namespace ccd;

extern "C" int ccd_ccl_label(const float* src, int mode, void* bits_u32, void* compact_u8, int* n_comp, int n_img, void* stream) {
  if (!src || n_img <= 0 || (mode != 0 && mode != 1) || (!bits_u32 && !compact_u8)) return CCD_ERR_ARG;
  static bool attr_set = false;
  const int smem = 3 * IMG_PX * (int)sizeof(int);
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(ccl_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  ccl_label_kernel<<<n_img, 256, smem, (cudaStream_t)stream>>>(src, mode, (unsigned*)bits_u32, (unsigned char*)compact_u8, n_comp);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_affine_theta(const double* m_inv, const int* src_hw, float* theta, int n, int img_h, int img_w, void* stream) {
  if (!m_inv || !src_hw || !theta || n <= 0 || img_h < 2 || img_w < 2) return CCD_ERR_ARG;
  affine_theta_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m_inv, src_hw, theta, n, img_h, img_w);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_kmeans_mask(const void* grey_u8, float* mask, int n_img, int H, int W, void* stream) {
  if (!grey_u8 || !mask || n_img <= 0 || H < 2 || W < 2 || (long long)H * W > (1 << 20)) return CCD_ERR_ARG;
  kmeans_mask_kernel<<<n_img, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)grey_u8, mask, H, W);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_warp_bits(const void* src_bits, const float* theta, void* dst_bits, int n_img, void* stream) {
  if (!src_bits || !theta || !dst_bits || n_img <= 0) return CCD_ERR_ARG;
  warp_bits_kernel<<<(n_img * IMG_PX + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const unsigned*)src_bits, theta,
                                                                                   (unsigned*)dst_bits, n_img);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_warp_mask(const float* src, const float* theta, float* dst, int n_img, void* stream) {
  if (!src || !theta || !dst || n_img <= 0) return CCD_ERR_ARG;
  warp_mask_kernel<<<(n_img * IMG_PX + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, theta, dst, n_img);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_bits_to_dense(const void* bits, float* dense, int n_img, void* stream) {
  if (!bits || !dense || n_img <= 0) return CCD_ERR_ARG;
  const size_t total = (size_t)n_img * SLOTS * IMG_PX;
  bits_to_dense_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const unsigned*)bits, dense, n_img);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
extern "C" int ccd_dense_to_bits(const float* dense, void* bits, int n_img, void* stream) {
  if (!bits || !dense || n_img <= 0) return CCD_ERR_ARG;
  dense_to_bits_kernel<<<(n_img * IMG_PX + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dense, (unsigned*)bits, n_img);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

// bits [2*n_view,32,128]; outputs tot4 [2*n_view,26], cnt [n_view], offs [n_view+1] (offs[n_view] = R), new_index [n_view,26]
extern "C" int ccd_char_plan(const void* bits, int* tot4, int* cnt, int* offs, void* new_index_u8, int n_view, void* stream) {
  if (!bits || !tot4 || !cnt || !offs || n_view <= 0) return CCD_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  char_index_kernel<<<2 * n_view, 256, 0, s>>>((const unsigned*)bits, tot4, cnt, n_view);
  CCD_LAUNCH_CHECK();
  char_offsets_kernel<<<1, 1024, 0, s>>>(cnt, offs, n_view);
  CCD_LAUNCH_CHECK();
  if (new_index_u8) {
    new_index_kernel<<<(n_view * SLOTS + 255) / 256, 256, 0, s>>>(cnt, (unsigned char*)new_index_u8, n_view);
    CCD_LAUNCH_CHECK();
  }
  return CCD_OK;
}

extern "C" int ccd_char_pool_fwd(const void* tokens, int tokens_bf16, const void* bits, const int* tot4, const int* cnt,
                                 const int* offs, float* rows, int n_view, int E, void* stream) {
  if (!tokens || !bits || !tot4 || !cnt || !offs || !rows || n_view <= 0 || E <= 0) return CCD_ERR_ARG;
  const int smem = SLOTS * E * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(char_pool_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 26 * 512 * 4));
    CCD_CUDA_CHECK(cudaFuncSetAttribute(char_pool_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 26 * 512 * 4));
    CCD_CUDA_CHECK(cudaFuncSetAttribute(char_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 26 * 512 * 4));
    attr_set = true;
  }
  if (E > 512) return CCD_ERR_ARG;
  if (tokens_bf16)
    char_pool_fwd_kernel<bf16><<<2 * n_view, 256, smem, (cudaStream_t)stream>>>((const bf16*)tokens, (const unsigned*)bits, tot4, cnt,
                                                                               offs, rows, n_view, E);
  else
    char_pool_fwd_kernel<float><<<2 * n_view, 256, smem, (cudaStream_t)stream>>>((const float*)tokens, (const unsigned*)bits, tot4,
                                                                                cnt, offs, rows, n_view, E);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}

extern "C" int ccd_char_pool_bwd(const float* drows, const void* bits, const int* tot4, const int* cnt, const int* offs,
                                 float* dtokens, int n_view, int E, void* stream) {
  if (!drows || !bits || !tot4 || !cnt || !offs || !dtokens || n_view <= 0 || E <= 0 || E > 512) return CCD_ERR_ARG;
  const int smem = SLOTS * E * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    CCD_CUDA_CHECK(cudaFuncSetAttribute(char_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 26 * 512 * 4));
    attr_set = true;
  }
  char_pool_bwd_kernel<<<2 * n_view, 256, smem, (cudaStream_t)stream>>>(drows, (const unsigned*)bits, tot4, cnt, offs, dtokens, n_view, E);
  CCD_LAUNCH_CHECK();
  return CCD_OK;
}
