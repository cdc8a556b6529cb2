// Pieces of the BatchNorm + ReLU + Dropout arithmetic shared by bn_act.cu and hei.cu: the counter-based
// dropout stream, the activation, and the kernels that turn per-CTA column partials into statistics.
// Everything is in an anonymous namespace: each translation unit gets its own copy.
#pragma once

#include "common.cuh"

namespace aread {
namespace {

constexpr int kBnThreads = 256;
constexpr int kStatCtas = kNumSMs * 4;   // upper bound on the number of column partials per reduction

// counter-based dropout stream: keep(element) is a pure function of (seed, salt, element index).  One 32-bit hash
// serves the two elements 2q and 2q + 1 (16 bits each), so kernels that walk consecutive elements pay half a hash
// per element; p is resolved to 1 / 65536.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// pairs at and beyond 2^32 (tensors of 2^33 elements and more): kept out of line so that the common path is one mix32
__device__ __noinline__ uint32_t dropout_inner_far(uint32_t hi, uint32_t ks) { return mix32(hi ^ ks); }
// `seed`, `salt` and a pair below 2^32 make the inner hash a per-launch constant that the compiler hoists out of the
// element loops: one mix32 per pair instead of two, the same bits as before.
__device__ __forceinline__ uint32_t dropout_pair_hash(uint64_t seed, uint32_t salt, uint64_t pair) {
  const uint32_t hi = static_cast<uint32_t>(pair >> 32);
  const uint32_t ks = salt ^ static_cast<uint32_t>(seed);
  uint32_t inner = mix32(ks);                          // loop-invariant
  if (hi != 0u) inner = dropout_inner_far(hi, ks);
  return mix32(static_cast<uint32_t>(pair) ^ inner ^ static_cast<uint32_t>(seed >> 32));
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t salt, uint64_t idx, uint32_t threshold) {
  const uint32_t h = dropout_pair_hash(seed, salt, idx >> 1);
  return ((idx & 1) ? (h >> 16) : (h & 0xffffu)) >= threshold;
}
// the dropout seed of a launch: read from device memory when the caller gave a slot (CUDA-graph replay: the
// launch parameters are frozen, the seed is not), else the by-value field
template <typename Args>
__device__ __forceinline__ uint64_t seed_of(const Args& a) {
  return a.seed_ptr != nullptr ? __ldg(a.seed_ptr) : a.seed;
}
inline uint32_t dropout_threshold(float p) {   // 16-bit scale; 0 = no dropout
  if (p <= 0.f) return 0u;
  const double t = static_cast<double>(p) * 65536.0;
  return t >= 65535.0 ? 0xffffu : static_cast<uint32_t>(t + 0.5);
}

// Sum of the CTA partials of one column pair, by 8 threads in interleaved order then in thread order.
// Block = 32 columns x 8; returns the totals to the threads with threadIdx.x < 32.
__device__ __forceinline__ void combine_partials(const float* __restrict__ partial, int n_partial, int width, int col,
                                                 float& s, float& q) {
  __shared__ float s_c[2][8][32];
  const int cx = threadIdx.x % 32, py = threadIdx.x / 32;
  float a = 0.f, b = 0.f;
  if (col < width) {
    for (int i = py; i < n_partial; i += 8) {
      a += partial[(static_cast<int64_t>(i) * 2 + 0) * width + col];
      b += partial[(static_cast<int64_t>(i) * 2 + 1) * width + col];
    }
  }
  s_c[0][py][cx] = a;
  s_c[1][py][cx] = b;
  __syncthreads();
  s = q = 0.f;
  if (py == 0) {
    for (int y = 0; y < 8; ++y) { s += s_c[0][y][cx]; q += s_c[1][y][cx]; }
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_finalize_kernel(const aread_bn_act_args a, const float* partial,
                                                               int n_partial) {
  const int col = blockIdx.x * 32 + threadIdx.x % 32;
  float s = 0.f, q = 0.f;
  if (a.training && !a.bn_skip) combine_partials(partial, n_partial, a.width, col, s, q);
  if (col >= a.width || threadIdx.x >= 32) return;
  float scale, shift;
  if (a.bn_skip) {
    scale = 1.f;
    shift = 0.f;
    a.mean[col] = 0.f;
    a.rstd[col] = 1.f;
  } else if (a.training) {
    const float inv_m = 1.f / static_cast<float>(a.m);
    const float d = s * inv_m;                       // mean - pivot
    const float mean = a.z[col] + d;                 // the pivot is the column's first row
    const float var = fmaxf(q * inv_m - d * d, 0.f);
    const float rstd = 1.f / sqrtf(var + a.eps);
    a.mean[col] = mean;
    a.rstd[col] = rstd;
    scale = a.gamma[col] * rstd;
    shift = a.beta[col] - mean * scale;
    const float unbiased = a.m > 1 ? var * (static_cast<float>(a.m) / static_cast<float>(a.m - 1)) : var;
    a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * mean;
    a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unbiased;
  } else {
    const float rstd = 1.f / sqrtf(a.running_var[col] + a.eps);
    a.mean[col] = a.running_mean[col];
    a.rstd[col] = rstd;
    scale = a.gamma[col] * rstd;
    shift = a.beta[col] - a.running_mean[col] * scale;
  }
  a.scale[col] = scale;
  a.shift[col] = shift;
}

__device__ __forceinline__ float act_value(float z, float scale, float shift, bool keep, float keep_scale) {
  const float y = fmaf(z, scale, shift);
  return (y > 0.f && keep) ? y * keep_scale : 0.f;
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_finalize_kernel(const aread_bn_act_bwd_args a, const float* partial,
                                                                   int n_partial, float* coef) {
  const int col = blockIdx.x * 32 + threadIdx.x % 32;
  float s1, s2;
  combine_partials(partial, n_partial, a.width, col, s1, s2);
  if (col >= a.width || threadIdx.x >= 32) return;
  if (a.bn_skip) {  // identity instead of BatchNorm: gamma / beta see no gradient, the bias sees sum(dy)
    if (a.d_gamma) a.d_gamma[col] = 0.f;
    if (a.d_beta) a.d_beta[col] = 0.f;
    if (a.d_bias) a.d_bias[col] = s1;
    coef[col] = 0.f;
    coef[a.width + col] = 0.f;
  } else {
    if (a.d_gamma) a.d_gamma[col] = s2;
    if (a.d_beta) a.d_beta[col] = s1;
    if (a.d_bias) a.d_bias[col] = 0.f;  // BatchNorm removes the column mean: the exact gradient is zero
    const float inv_m = 1.f / static_cast<float>(a.m);
    coef[col] = s1 * inv_m;
    coef[a.width + col] = s2 * inv_m;
  }
}


}  // namespace
}  // namespace aread
