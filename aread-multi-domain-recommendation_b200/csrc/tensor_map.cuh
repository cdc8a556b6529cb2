// Host-side helpers shared by the tcgen05 kernels: CUtensorMap encoding through the driver entry point (no libcuda
// link dependency), per-device kernel attribute bookkeeping and the SM count of the current device.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace aread {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols], 128B swizzle
inline int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_cols = 64) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AREAD_OK;
}

// fp32 row-major [rows, cols] output, box = [32 rows, 32 cols] (one 128-byte swizzle row per matrix row)
inline int make_store_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled (store) failed with CUresult %d", (int)r);
  return AREAD_OK;
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
inline bool first_use_on_device(uint64_t* seen) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = uint64_t{1} << (dev & 63);
  if (*seen & bit) return false;
  *seen |= bit;
  return true;
}

inline int sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = sms[dev & 63];
  if (n == 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = kNumSMs;
  return n;
}

}  // namespace aread
