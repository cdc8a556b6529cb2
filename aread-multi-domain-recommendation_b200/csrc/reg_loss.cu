// L2 regulariser over a list of parameter tensors in one launch:  sum_i l2_i * sum(w_i^2), and its
// gradient 2 * l2_i * w_i.  Replaces BaseModel.get_regularization_loss (model/layer.py:96-112), which
// the trainer evaluates every step over the whole embedding table and ~110 weight tensors
// (run.py:599, 644, 677).  HBM-bound: the table is read once.
//
// Work is cut into fixed 4096-element chunks over the concatenation of all tensors; chunk c always
// covers the same elements and the chunk partials are added in chunk order, so the result is
// bit-reproducible run to run.
#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;

// chunk -> (tensor, offset): binary search over the exclusive prefix of chunk counts
__device__ __forceinline__ int find_tensor(const int64_t* __restrict__ chunk_start, int n, int64_t chunk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads) l2_partial_kernel(const aread_l2_reg_args a, float* __restrict__ partial) {
  __shared__ float s_red[kThreads / 32];
  for (int64_t chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x) {
    const int t = find_tensor(a.chunk_start, a.n_tensors, chunk);
    const float* __restrict__ w = a.tensors[t];
    const int64_t begin = (chunk - a.chunk_start[t]) * kChunk;
    const int64_t end = min(a.sizes[t], begin + kChunk);
    float acc = 0.f;
    for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) {
      const float v = __ldg(w + i);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) s_red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < kThreads / 32; ++i) tot += s_red[i];
      partial[chunk] = tot * a.l2[t];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) l2_final_kernel(int64_t n_chunks, const float* __restrict__ partial,
                                                        float* __restrict__ out) {
  // fixed-shape tree: 1024 strided serial sums, then a shared-memory tree
  __shared__ float s[1024];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n_chunks; i += 1024) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

// grad_i = (2 * l2_i * g) * w_i, written to grads[i]
__global__ void __launch_bounds__(kThreads) l2_grad_kernel(const aread_l2_reg_args a, const float* __restrict__ g_out) {
  const float g = g_out ? __ldg(g_out) : 1.f;
  for (int64_t chunk = blockIdx.x; chunk < a.n_chunks; chunk += gridDim.x) {
    const int t = find_tensor(a.chunk_start, a.n_tensors, chunk);
    const float* __restrict__ w = a.tensors[t];
    float* __restrict__ d = a.grads[t];
    const float c = 2.f * a.l2[t] * g;
    const int64_t begin = (chunk - a.chunk_start[t]) * kChunk;
    const int64_t end = min(a.sizes[t], begin + kChunk);
    for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) d[i] = c * __ldg(w + i);
  }
}

}  // namespace
}  // namespace aread

extern "C" {

int64_t aread_l2_reg_chunk(void) { return aread::kChunk; }

int aread_l2_reg_fwd(const aread_l2_reg_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "l2_reg_fwd: null args");
  const aread_l2_reg_args& a = *args;
  AREAD_REQUIRE(a.n_tensors >= 0 && a.n_chunks >= 0 && a.out && a.workspace, "l2_reg_fwd: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partial = static_cast<float*>(a.workspace);
  if (a.n_chunks > 0) {
    AREAD_REQUIRE(a.tensors && a.sizes && a.l2 && a.chunk_start, "l2_reg_fwd: null pointer");
    AREAD_REQUIRE(a.workspace_bytes >= static_cast<size_t>(a.n_chunks) * 4, "l2_reg_fwd: workspace too small");
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    AREAD_LAUNCH(l2_partial_kernel, static_cast<unsigned>(a.n_chunks < cap ? a.n_chunks : cap), kThreads, 0, stream, a,
                 partial);
  }
  AREAD_LAUNCH(l2_final_kernel, 1, 1024, 0, stream, a.n_chunks, partial, a.out);
  return AREAD_OK;
}

int aread_l2_reg_bwd(const aread_l2_reg_args* args, const float* g_out, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "l2_reg_bwd: null args");
  const aread_l2_reg_args& a = *args;
  if (a.n_chunks <= 0) return AREAD_OK;
  AREAD_REQUIRE(a.tensors && a.grads && a.sizes && a.l2 && a.chunk_start, "l2_reg_bwd: null pointer");
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  AREAD_LAUNCH(l2_grad_kernel, static_cast<unsigned>(a.n_chunks < cap ? a.n_chunks : cap), kThreads, 0,
               static_cast<cudaStream_t>(stream_), a, g_out);
  return AREAD_OK;
}

}  // extern "C"
