// Host-side TMA descriptor (CUtensorMap) construction + cache.  The driver entry point is fetched through the
// runtime (cudaGetDriverEntryPoint) so the shared library does not link libcuda directly.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <unordered_map>

namespace ccd {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

struct TmapKey {
  uint64_t ptr, rows, cols, ld;
  uint32_t box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.ld + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (((uint64_t)k.box_rows << 32 | k.box_cols) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = box_rows x box_cols,
// box_cols * 2 bytes must be 128 (SWIZZLE_128B).  Out-of-bounds elements are zero-filled.
// Returns false on failure.
inline bool get_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{(uint64_t)ptr, rows, cols, ld, box_rows, box_cols};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  if (((uint64_t)ptr & 15) || ((ld * 2) & 15) || box_cols * 2 != 128 || box_rows > 256) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return true;
}

// 5-D bf16 view {C (inner, contiguous), W, P, H, N} of an NHWC-like activation for implicit-GEMM convolutions:
// strides in ELEMENTS for dims 1..4; box = {64 channels, box_w, 1, box_h, 1} (128B swizzle, zero fill out of bounds:
// that is the convolution padding).  Not cached (cheap, a handful per step).
inline bool make_tmap_bf16_5d(CUtensorMap* out, const void* ptr, const uint64_t dims[5], const uint64_t strides_elems[4],
                              uint32_t box_w, uint32_t box_h) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  if ((uint64_t)ptr & 15) return false;
  cuuint64_t gdim[5] = {dims[0], dims[1], dims[2], dims[3], dims[4]};
  cuuint64_t gstride[4] = {strides_elems[0] * 2, strides_elems[1] * 2, strides_elems[2] * 2, strides_elems[3] * 2};
  for (int i = 0; i < 4; ++i)
    if (gstride[i] & 15) return false;
  cuuint32_t box[5] = {64, box_w, 1, box_h, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace ccd
