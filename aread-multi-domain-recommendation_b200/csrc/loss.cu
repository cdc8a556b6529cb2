// Bagging loss of the HEI heads in one pass:  loss = (1 / T) * sum_t mean_b BCE(p[t, b], y[b]), and its
// gradient w.r.t. the probabilities.  Replaces the trainer's `sum(BCELoss(y_stack[t], y)) / n_act`
// (run.py:643-644, 672-677; BCELoss at run.py:833), i.e. T small kernels + T autograd nodes.
// Element arithmetic is torch's binary_cross_entropy: logs clamped at -100, backward denominator
// clamped at 1e-12.  HBM-bound: reads T*m probabilities + m labels, writes T*m gradients.
//
// Reduction: tower t is cut into fixed 4096-sample chunks; the chunk partials are added in (t, chunk)
// order by one CTA, so the value is bit-reproducible.
#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;

__global__ void __launch_bounds__(kThreads) bce_partial_kernel(const aread_bagging_bce_args a, float* __restrict__ partial,
                                                               int n_chunks) {
  __shared__ float s_red[kThreads / 32];
  const int t = blockIdx.y;
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * kChunk;
  const int64_t end = begin + kChunk < a.m ? begin + kChunk : a.m;
  const float* __restrict__ p = a.probs + static_cast<int64_t>(t) * a.m;
  float* __restrict__ d = a.d_probs ? a.d_probs + static_cast<int64_t>(t) * a.m : nullptr;
  const float scale = 1.f / (static_cast<float>(a.m) * static_cast<float>(a.n_tower));
  float acc = 0.f;
  for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) {
    const float pi = __ldg(p + i), yi = __ldg(a.labels + i);
    const float lp = fmaxf(logf(pi), -100.f), lq = fmaxf(log1pf(-pi), -100.f);
    acc += (yi - 1.f) * lq - yi * lp;
    if (d) d[i] = scale * ((pi - yi) / fmaxf((1.f - pi) * pi, 1e-12f));
  }
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) s_red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) tot += s_red[i];
    partial[static_cast<int64_t>(t) * n_chunks + blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(256) bce_final_kernel(int64_t n, const float* __restrict__ partial, float scale,
                                                        float* __restrict__ out) {
  __shared__ float s[256];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0] * scale;
}

}  // namespace
}  // namespace aread

extern "C" {

size_t aread_bagging_bce_workspace_bytes(int64_t m, int32_t n_tower) {
  const int64_t chunks = (m + aread::kChunk - 1) / aread::kChunk;
  return static_cast<size_t>((chunks > 0 ? chunks : 1) * (n_tower > 0 ? n_tower : 1)) * sizeof(float);
}

int aread_bagging_bce(const aread_bagging_bce_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "bagging_bce: null args");
  const aread_bagging_bce_args& a = *args;
  AREAD_REQUIRE(a.m > 0 && a.n_tower > 0 && a.n_tower <= 65535, "bagging_bce: bad shape m=%lld n_tower=%d",
                (long long)a.m, a.n_tower);
  AREAD_REQUIRE(a.probs && a.labels && a.loss && a.workspace, "bagging_bce: null pointer");
  AREAD_REQUIRE(a.workspace_bytes >= aread_bagging_bce_workspace_bytes(a.m, a.n_tower), "bagging_bce: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int chunks = ceil_div(a.m, kChunk);
  float* partial = static_cast<float*>(a.workspace);
  AREAD_LAUNCH(bce_partial_kernel, dim3(chunks, a.n_tower), kThreads, 0, stream, a, partial, chunks);
  AREAD_LAUNCH(bce_final_kernel, 1, 256, 0, stream, static_cast<int64_t>(chunks) * a.n_tower, partial,
               1.f / (static_cast<float>(a.m) * static_cast<float>(a.n_tower)), a.loss);
  return AREAD_OK;
}

}  // extern "C"
