// Host-side helper: CUtensorMap of an activation image for tensor-map TMA (cp.async.bulk.tensor).  The encoder is looked
// up through the runtime (cudaGetDriverEntryPoint), so the library does not link against libcuda.
#pragma once

#include <cuda.h>

#include <unordered_map>

#include "common.cuh"

namespace mgb {

// Tensor map of an activation image [nchunks][Rp][8] bf16, seen as a 2-D array of 8-byte words: dim0 = 2 words per row
// (contiguous, Rp rows), dim1 = chunks.  A box of 256 words x k chunks is 128 rows x 8k channels and lands in shared memory
// as [chunk][128 rows][16 B] — the no-swizzle core-matrix order both GEMM families read.  Coordinates: (2 * row, chunk).
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<TmapEncodeFn>(p);
  }();
  return fn;
}
// Encoding costs ~1 us of host time and a training step needs several hundred maps over a few dozen distinct images whose
// addresses repeat from step to step (workspace, and the caching allocator's activation stash), so maps are memoised.
struct TmapKey {
  const void* img; int nchunks, Rp, box;
  bool operator==(const TmapKey& o) const { return img == o.img && nchunks == o.nchunks && Rp == o.Rp && box == o.box; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.img) >> 4;
    h = h * 1000003u ^ (size_t)k.nchunks;
    h = h * 1000003u ^ (size_t)k.Rp;
    return h * 1000003u ^ (size_t)k.box;
  }
};
inline int make_image_map(CUtensorMap* m, const void* img, int nchunks, int Rp, int box_chunks) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{img, nchunks, Rp, box_chunks};
  auto it = cache.find(key);
  if (it != cache.end()) { *m = it->second; return MGB_OK; }
  if (cache.size() > 8192) cache.clear();
  TmapEncodeFn fn = tmap_encode_fn();
  MGB_REQUIRE(fn != nullptr, MGB_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)Rp * 2, (cuuint64_t)nchunks};
  const cuuint64_t strides[1] = {(cuuint64_t)Rp * 16};
  const cuuint32_t box[2] = {256, (cuuint32_t)box_chunks};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(img), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MGB_REQUIRE(r == CUDA_SUCCESS, MGB_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for an image of %d chunks x %d rows", (int)r,
              nchunks, Rp);
  cache.emplace(key, *m);
  return MGB_OK;
}

// The same image as a 3-D array of 8-byte words whose innermost line is EIGHT rows (16 words = 128 B, one L2 line): dim0 =
// the 16 words of a group of 8 rows, dim1 = row groups, dim2 = chunks.  A box of 16 x (box_rows / 8) x box_chunks lands as
// [chunk][box_rows][16 B] like the 2-D form, but box_rows may exceed 128 (a 128-row tile plus its convolution halo, a
// multiple of 8) and row groups outside the image are zero-filled.  Coordinates: (0, row / 8, chunk).  (A 16-byte
// innermost line - one row - made the TMA unit walk 1024 lines per box and cost 2.6x in load throughput.)
inline int make_image_map3(CUtensorMap* m, const void* img, int nchunks, int Rp, int box_rows, int box_chunks) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{img, nchunks, Rp, box_rows * 1024 + box_chunks};
  auto it = cache.find(key);
  if (it != cache.end()) { *m = it->second; return MGB_OK; }
  if (cache.size() > 8192) cache.clear();
  TmapEncodeFn fn = tmap_encode_fn();
  MGB_REQUIRE(fn != nullptr, MGB_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  MGB_REQUIRE(box_rows >= 8 && box_rows <= 2048 && box_rows % 8 == 0 && Rp % 8 == 0, MGB_E_ARG, "TMA box of %d rows over %d", box_rows, Rp);
  const cuuint64_t dims[3] = {16, (cuuint64_t)Rp / 8, (cuuint64_t)nchunks};
  const cuuint64_t strides[2] = {128, (cuuint64_t)Rp * 16};
  const cuuint32_t box[3] = {16, (cuuint32_t)box_rows / 8, (cuuint32_t)box_chunks};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(img), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MGB_REQUIRE(r == CUDA_SUCCESS, MGB_E_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d) for an image of %d chunks x %d rows", (int)r,
              nchunks, Rp);
  cache.emplace(key, *m);
  return MGB_OK;
}

}  // namespace mgb
