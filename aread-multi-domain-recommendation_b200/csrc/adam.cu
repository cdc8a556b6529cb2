// Multi-tensor Adam step in one launch (coupled weight decay, the trainer's settings run.py:830-831).
// Replaces torch.optim.Adam.step over the ~460 parameter tensors of AREAD: one pass that reads
// p, g, m, v and writes p, m, v (7 x 4 B per element) instead of ~20 foreach passes.  Per-tensor step
// counts arrive as precomputed bias corrections, so tensors whose gradient is None this step are
// simply not in the list (their moments and step count stay untouched, like the reference).
//
// Arithmetic per element, in torch's order (torch/optim/adam.py _single_tensor_adam):
//   g' = g + wd * p ;  m += (g' - m) * (1 - b1) ;  v = v * b2 + (1 - b2) * g' * g'
//   p -= step_size * m / (sqrt(v) / bc2_sqrt + eps),  step_size = lr / (1 - b1^t), bc2_sqrt = sqrt(1 - b2^t)
#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;

__device__ __forceinline__ int find_tensor(const int64_t* __restrict__ chunk_start, int n, int64_t chunk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads) adam_kernel(const aread_adam_args a) {
  for (int64_t chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x) {
    const int t = find_tensor(a.chunk_start, a.n_tensors, chunk);
    float* __restrict__ p = a.params[t];
    const float* __restrict__ g = a.grads[t];
    float* __restrict__ m = a.exp_avg[t];
    float* __restrict__ v = a.exp_avg_sq[t];
    float step_size, bc2_sqrt;
    if (a.step_counts != nullptr) {      // counters were advanced by adam_tick_kernel, launched just before
      const double n = static_cast<double>(a.step_counts[a.slot[t]]);
      step_size = static_cast<float>(a.lr / (1.0 - pow(a.beta1_d, n)));
      bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(a.beta2_d, n)));
    } else {
      step_size = a.step_size[t];
      bc2_sqrt = a.bc2_sqrt[t];
    }
    const float c = a.l2_twice ? a.l2_twice[t] : 0.f;
    const int64_t begin = (chunk - a.chunk_start[t]) * kChunk;
    const int64_t end = min(a.sizes[t], begin + kChunk);
    const float wd = a.weight_decay, b1 = a.beta1, b2 = a.beta2, eps = a.eps;
    auto update = [&](float& pi, float gi, float& mi, float& vi) {
      if (c != 0.f) gi = __fadd_rn(gi, __fmul_rn(c, pi));   // rounded like the separate regulariser gradient + add
      gi = gi + wd * pi;
      mi = mi + (gi - mi) * (1.f - b1);
      vi = vi * b2 + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pi = pi - step_size * (mi / denom);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(g)) & 15) == 0;          // (chunks start at multiples of 4096 elements)
    int64_t i0 = begin;
    if (vec) {
      const int64_t n4 = (end - begin) / 4;
      for (int64_t q = threadIdx.x; q < n4; q += kThreads) {
        const int64_t i = begin + q * 4;
        float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i),
               v4 = *reinterpret_cast<float4*>(v + i);
        const float4 g4 = g ? __ldg(reinterpret_cast<const float4*>(g + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        update(p4.x, g4.x, m4.x, v4.x);
        update(p4.y, g4.y, m4.y, v4.y);
        update(p4.z, g4.z, m4.z, v4.z);
        update(p4.w, g4.w, m4.w, v4.w);
        *reinterpret_cast<float4*>(p + i) = p4;
        *reinterpret_cast<float4*>(m + i) = m4;
        *reinterpret_cast<float4*>(v + i) = v4;
      }
      i0 = begin + n4 * 4;
    }
    for (int64_t i = i0 + threadIdx.x; i < end; i += kThreads) {
      float pi = p[i], mi = m[i], vi = v[i];
      update(pi, g ? __ldg(g + i) : 0.f, mi, vi);
      p[i] = pi;
      m[i] = mi;
      v[i] = vi;
    }
  }
}

__global__ void __launch_bounds__(kThreads) adam_tick_kernel(float* __restrict__ step_counts,
                                                             const int64_t* __restrict__ slot, int n) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) step_counts[slot[t]] += 1.f;
}

constexpr int64_t kCopyChunk = 64 * 1024;   // bytes

// dst[t] = src[t] (or zero) for a list of tensors, 16-byte accesses where the pair is aligned
__global__ void __launch_bounds__(kThreads) multi_copy_kernel(const aread_multi_copy_args a) {
  for (int64_t chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x) {
    const int t = find_tensor(a.chunk_start, a.n_tensors, chunk);
    char* __restrict__ d = static_cast<char*>(a.dst[t]);
    const char* __restrict__ s = a.src ? static_cast<const char*>(a.src[t]) : nullptr;
    const int64_t begin = (chunk - a.chunk_start[t]) * kCopyChunk;
    const int64_t end = min(a.bytes[t], begin + kCopyChunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
    int64_t i0 = begin;
    if (vec) {
      const int64_t n16 = (end - begin) / 16;
      for (int64_t q = threadIdx.x; q < n16; q += kThreads) {
        const int64_t i = begin + q * 16;
        *reinterpret_cast<uint4*>(d + i) = s ? *reinterpret_cast<const uint4*>(s + i) : make_uint4(0, 0, 0, 0);
      }
      i0 = begin + n16 * 16;
    }
    for (int64_t i = i0 + threadIdx.x * 4; i < end; i += kThreads * 4)
      *reinterpret_cast<uint32_t*>(d + i) = s ? *reinterpret_cast<const uint32_t*>(s + i) : 0u;
  }
}

// dst[t] = bf16(src[t]); bytes[] counts SOURCE bytes (fp32), 4 elements per thread and iteration
__global__ void __launch_bounds__(kThreads) multi_cast_bf16_kernel(const aread_multi_copy_args a) {
  for (int64_t chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x) {
    const int t = find_tensor(a.chunk_start, a.n_tensors, chunk);
    const float* __restrict__ s = static_cast<const float*>(a.src[t]);
    __nv_bfloat16* __restrict__ d = static_cast<__nv_bfloat16*>(a.dst[t]);
    const int64_t begin = (chunk - a.chunk_start[t]) * (kCopyChunk / 4);
    const int64_t end = min(a.bytes[t] / 4, begin + kCopyChunk / 4);
    for (int64_t i = begin + threadIdx.x * 4; i < end; i += kThreads * 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(s + i));
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(d + i) = pk;
    }
  }
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_multi_cast_bf16(const aread_multi_copy_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "multi_cast_bf16: null args");
  const aread_multi_copy_args& a = *args;
  if (a.n_tensors <= 0 || a.n_chunks <= 0) return AREAD_OK;
  AREAD_REQUIRE(a.dst && a.src && a.bytes && a.chunk_start, "multi_cast_bf16: null pointer");
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  AREAD_LAUNCH(multi_cast_bf16_kernel, static_cast<unsigned>(a.n_chunks < cap ? a.n_chunks : cap), kThreads, 0,
               static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

int64_t aread_multi_copy_chunk(void) { return aread::kCopyChunk; }

int aread_multi_copy(const aread_multi_copy_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "multi_copy: null args");
  const aread_multi_copy_args& a = *args;
  if (a.n_tensors <= 0 || a.n_chunks <= 0) return AREAD_OK;
  AREAD_REQUIRE(a.dst && a.bytes && a.chunk_start, "multi_copy: null pointer");
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  AREAD_LAUNCH(multi_copy_kernel, static_cast<unsigned>(a.n_chunks < cap ? a.n_chunks : cap), kThreads, 0,
               static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

int64_t aread_adam_chunk(void) { return aread::kChunk; }

int aread_adam_step(const aread_adam_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "adam_step: null args");
  const aread_adam_args& a = *args;
  if (a.n_tensors <= 0 || a.n_chunks <= 0) return AREAD_OK;
  AREAD_REQUIRE(a.params && a.grads && a.exp_avg && a.exp_avg_sq && a.sizes && a.chunk_start, "adam_step: null pointer");
  AREAD_REQUIRE(a.step_counts ? a.slot != nullptr : (a.step_size && a.bc2_sqrt), "adam_step: null step information");
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  if (a.step_counts != nullptr)   // tensors of one launch are distinct parameters: no two threads share a counter
    AREAD_LAUNCH(adam_tick_kernel, (a.n_tensors + kThreads - 1) / kThreads, kThreads, 0,
                 static_cast<cudaStream_t>(stream_), a.step_counts, a.slot, a.n_tensors);
  AREAD_LAUNCH(adam_kernel, static_cast<unsigned>(a.n_chunks < cap ? a.n_chunks : cap), kThreads, 0,
               static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

}  // extern "C"
