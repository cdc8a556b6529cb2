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

// ---------------------------------------------------------------------------------------- output heads
// p[t, b] = sigmoid(head_cross[b, t] + <h[b, t, :], w_tail[t, :]> + lin[b])      (model/aread.py:307-310; the part
// of towers_linear that multiplies cn_out arrives pre-reduced in head_cross, see rowpass.cu)
// One thread per sample; the transposed [T, m] output is written coalesced (threads run along b).
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(const aread_head_args a) {
  const int T = a.n_tower, W = a.width;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < a.m;
       b += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float lin = __ldg(a.lin + b);
    const float* h = a.h + b * T * W;
    for (int t = 0; t < T; ++t) {
      float acc = 0.f;
      for (int c = 0; c < W; ++c) acc = fmaf(__ldg(h + t * W + c), __ldg(a.w_tail + t * W + c), acc);
      const float z = __ldg(a.head_cross + b * T + t) + acc + lin;
      a.probs[static_cast<int64_t>(t) * a.m + b] = 1.f / (1.f + expf(-z));
    }
  }
}

// dz[b, t] = d_p[t, b] * p (1 - p);  d_lin[b] = sum_t dz;  d_h[b, t, :] = dz * w_tail[t, :];
// d_w_tail[t, :] = sum_b dz[b, t] * h[b, t, :]  -- warp sums into per-warp shared accumulators, warps added in
// order, one partial per CTA, CTAs added in order by head_wgrad_reduce_kernel (bit-reproducible).
constexpr int kHeadCtas = kNumSMs * 2;

__global__ void __launch_bounds__(kThreads) head_bwd_kernel(const aread_head_args a, float* __restrict__ partial) {
  extern __shared__ float s_acc[];  // [warps][T * W]
  const int T = a.n_tower, W = a.width, TW = T * W;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (kThreads / 32) * TW; i += kThreads) s_acc[i] = 0.f;
  __syncthreads();
  float* acc = s_acc + warp * TW;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * kThreads; base < a.m; base += stride) {
    const int64_t b = base + threadIdx.x;
    const bool valid = b < a.m;
    float dl = 0.f;
    for (int t = 0; t < T; ++t) {
      float dz = 0.f;
      if (valid) {
        const float p = __ldg(a.probs + static_cast<int64_t>(t) * a.m + b);
        dz = __ldg(a.d_probs + static_cast<int64_t>(t) * a.m + b) * p * (1.f - p);
        a.dz[b * T + t] = dz;
        dl += dz;
      }
      for (int c = 0; c < W; ++c) {
        float g = 0.f;
        if (valid) {
          a.d_h[(b * T + t) * W + c] = dz * __ldg(a.w_tail + t * W + c);
          g = dz * __ldg(a.h + (b * T + t) * W + c);
        }
        if (a.d_w_tail != nullptr) {
          g = warp_sum(g);
          if (lane == 0) acc[t * W + c] += g;
        }
      }
    }
    if (valid) a.d_lin[b] = dl;
  }
  if (a.d_w_tail == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < TW; i += kThreads) {
    float v = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) v += s_acc[w * TW + i];
    partial[static_cast<int64_t>(blockIdx.x) * TW + i] = v;
  }
}

// The same for heads of width 8 and up to 12 towers (every shipped configuration): 128-bit loads / stores of the rows
// and the head weight gradient accumulated per thread in registers over all of its rows -- one warp reduction per
// (tower, column) at the end instead of one per row block.  Rows are added in thread order, lanes by the butterfly,
// warps and CTAs in order: bit-reproducible.
constexpr int kHeadMaxT = 12;
__global__ void __launch_bounds__(kThreads) head_bwd_w8_kernel(const aread_head_args a, float* __restrict__ partial) {
  extern __shared__ float s_acc[];  // [warps][T * 8]
  const int T = a.n_tower, TW = T * 8;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  float acc[kHeadMaxT][8];
#pragma unroll
  for (int t = 0; t < kHeadMaxT; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; b < a.m; b += stride) {
    float dl = 0.f;
#pragma unroll
    for (int t = 0; t < kHeadMaxT; ++t) {
      if (t < T) {
        const float p = __ldg(a.probs + static_cast<int64_t>(t) * a.m + b);
        const float dz = __ldg(a.d_probs + static_cast<int64_t>(t) * a.m + b) * p * (1.f - p);
        a.dz[b * T + t] = dz;
        dl += dz;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w_tail + t * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w_tail + t * 8 + 4));
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(a.h + (b * T + t) * 8));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(a.h + (b * T + t) * 8 + 4));
        float4* dh = reinterpret_cast<float4*>(a.d_h + (b * T + t) * 8);
        dh[0] = make_float4(dz * w0.x, dz * w0.y, dz * w0.z, dz * w0.w);
        dh[1] = make_float4(dz * w1.x, dz * w1.y, dz * w1.z, dz * w1.w);
        acc[t][0] = fmaf(dz, h0.x, acc[t][0]); acc[t][1] = fmaf(dz, h0.y, acc[t][1]);
        acc[t][2] = fmaf(dz, h0.z, acc[t][2]); acc[t][3] = fmaf(dz, h0.w, acc[t][3]);
        acc[t][4] = fmaf(dz, h1.x, acc[t][4]); acc[t][5] = fmaf(dz, h1.y, acc[t][5]);
        acc[t][6] = fmaf(dz, h1.z, acc[t][6]); acc[t][7] = fmaf(dz, h1.w, acc[t][7]);
      }
    }
    a.d_lin[b] = dl;
  }
  if (a.d_w_tail == nullptr) return;
#pragma unroll
  for (int t = 0; t < kHeadMaxT; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (t < T) {
        const float v = warp_sum(acc[t][c]);
        if (lane == 0) s_acc[warp * TW + t * 8 + c] = v;
      }
  __syncthreads();
  for (int i = threadIdx.x; i < TW; i += kThreads) {
    float v = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) v += s_acc[w * TW + i];
    partial[static_cast<int64_t>(blockIdx.x) * TW + i] = v;
  }
}

__global__ void __launch_bounds__(kThreads) head_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial,
                                                                     int n, float* __restrict__ out) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  for (int p = 0; p < n_partial; ++p) v += partial[static_cast<int64_t>(p) * n + i];
  out[i] = v;
}

}  // namespace
}  // namespace aread

extern "C" {

size_t aread_head_workspace_bytes(int32_t n_tower, int32_t width) {
  return aread::align_up(sizeof(float) * static_cast<size_t>(aread::kHeadCtas) * (n_tower > 0 ? n_tower : 1) *
                             (width > 0 ? width : 1),
                         256);
}

int aread_head(const aread_head_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "head: null args");
  const aread_head_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.n_tower > 0 && a.width > 0, "head: bad shape");
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.h && a.w_tail && a.probs, "head: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t g = (a.m + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  const unsigned grid = static_cast<unsigned>(g > cap ? cap : g);
  if (a.d_probs == nullptr) {
    AREAD_REQUIRE(a.head_cross && a.lin, "head: null forward input");
    AREAD_LAUNCH(head_fwd_kernel, grid, kThreads, 0, stream, a);
  } else {
    AREAD_REQUIRE(a.dz && a.d_lin && a.d_h, "head: null gradient output");
    const int tw = a.n_tower * a.width;
    const unsigned bgrid = grid < static_cast<unsigned>(kHeadCtas) ? grid : static_cast<unsigned>(kHeadCtas);
    const size_t smem = sizeof(float) * (kThreads / 32) * tw;
    AREAD_REQUIRE(smem <= 48 * 1024, "head: %d x %d head weights do not fit the reduction buffer", a.n_tower, a.width);
    AREAD_REQUIRE(a.d_w_tail == nullptr || (a.workspace && a.workspace_bytes >= aread_head_workspace_bytes(a.n_tower, a.width)),
                  "head: workspace too small");
    float* partial = static_cast<float*>(a.workspace);
    const bool w8 = a.width == 8 && a.n_tower <= kHeadMaxT && reinterpret_cast<uintptr_t>(a.h) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(a.d_h) % 16 == 0 && reinterpret_cast<uintptr_t>(a.w_tail) % 16 == 0;
    if (w8) {
      AREAD_LAUNCH(head_bwd_w8_kernel, bgrid, kThreads, smem, stream, a, partial);
    } else {
      AREAD_LAUNCH(head_bwd_kernel, bgrid, kThreads, smem, stream, a, partial);
    }
    if (a.d_w_tail != nullptr)
      AREAD_LAUNCH(head_wgrad_reduce_kernel, ceil_div(tw, kThreads), kThreads, 0, stream, partial, static_cast<int>(bgrid),
                   tw, a.d_w_tail);
  }
  return AREAD_OK;
}

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
