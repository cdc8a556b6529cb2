// Shared plumbing of libaread_sm100.so: error reporting, launch accounting, small device helpers.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/aread_sm100.h"

namespace aread {

char* last_error_buf();              // thread-local, 512 bytes
std::atomic<uint64_t>& launch_counter();

inline int fail(aread_status code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return static_cast<int>(code);
}

#define AREAD_REQUIRE(cond, ...)                                     \
  do {                                                               \
    if (!(cond)) return ::aread::fail(AREAD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define AREAD_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return ::aread::fail(AREAD_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                  \
                           cudaGetErrorString(err__), __FILE__, __LINE__);                 \
  } while (0)

// Every kernel launch of the library goes through this so that aread_launch_count() is exact.
#define AREAD_LAUNCH(kernel, grid, block, smem, stream, ...)                               \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                            \
    ::aread::launch_counter().fetch_add(1, std::memory_order_relaxed);                     \
    AREAD_CUDA(cudaGetLastError());                                                        \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace aread
