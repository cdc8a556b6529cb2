// Everything in the AREAD trunk that is a dot product of the flattened embedding row X[b, :] with
// a parameter vector, as ONE skinny fp32 product P = X . Wcat^T plus a per-row scalar epilogue:
//
//   column 0                linear term            w_lin . X                 (model/layer.py:122-126)
//   next n_gate * n_expert  MMoE gate logits       W_g[e] . X                (model/aread.py:152)
//   next n_cross            cross-network dots     w_k . X                   (model/layer.py:533-537)
//   next n_head             output heads           w_out_t[:E] . X           (model/aread.py:307)
//
// The cross network never materialises: with c_0 = X, c_{k+1} = X (w_k . c_k) + b_k + c_k every
// c_k has the form alpha_k X + beta_k (alpha per row, beta_k = b_0 + .. + b_{k-1} shared), so
//   s_k = w_k . c_k = alpha_k (w_k . X) + w_k . beta_k ,  alpha_{k+1} = alpha_k + s_k ,
// and the head's cross part is w_out_t[:E] . c_n = alpha_n (w_out_t[:E] . X) + w_out_t[:E] . beta_n.
// The row-independent constants (biases, w_k . beta_k, w_out_t . beta_n) arrive in `offset[j]`.
// fp32 throughout (these terms feed the logit directly).
#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerWarp = 4;
constexpr int kJChunk = 32;

// P[b, j] = sum_e X[b, e] * W[j, e]; one warp per 4 rows, lanes over e, W staged in shared memory
__global__ void __launch_bounds__(kThreads) rowdots_fwd_kernel(int64_t m, int E, int nj, const float* __restrict__ x,
                                                               const float* __restrict__ w, float* __restrict__ p,
                                                               int ldp) {
  extern __shared__ float s_w[];  // [jc][E]
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int warps = blockDim.x / 32;
  for (int j0 = 0; j0 < nj; j0 += kJChunk) {
    const int jc = min(kJChunk, nj - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < jc * E; i += blockDim.x) s_w[i] = w[static_cast<int64_t>(j0) * E + i];
    __syncthreads();
    for (int64_t b0 = (static_cast<int64_t>(blockIdx.x) * warps + warp) * kRowsPerWarp; b0 < m;
         b0 += static_cast<int64_t>(gridDim.x) * warps * kRowsPerWarp) {
      float acc[kRowsPerWarp][kJChunk];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
        for (int j = 0; j < kJChunk; ++j) acc[r][j] = 0.f;
      for (int e = lane; e < E; e += 32) {
        float xv[kRowsPerWarp];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) xv[r] = b0 + r < m ? __ldg(x + (b0 + r) * E + e) : 0.f;
#pragma unroll
        for (int j = 0; j < kJChunk; ++j) {
          if (j < jc) {
            const float wv = s_w[j * E + e];
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) acc[r][j] = fmaf(xv[r], wv, acc[r][j]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
#pragma unroll
        for (int j = 0; j < kJChunk; ++j) {
          if (j < jc) {
            const float v = warp_sum(acc[r][j]);
            if (lane == 0 && b0 + r < m) p[(b0 + r) * ldp + j0 + j] = v;
          }
        }
      }
    }
  }
}

// per-row scalar epilogue, one thread per row
__global__ void __launch_bounds__(kThreads) rowpass_epilogue_kernel(const aread_rowpass_args a) {
  const int ng = a.n_gate, ne = a.n_expert, nc = a.n_cross, nh = a.n_head;
  const int c_gate = 1, c_cross = 1 + ng * ne, c_head = c_cross + nc;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; b < a.m;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* p = a.p + b * a.ldp;
    a.lin[b] = p[0] + a.offset[0];
    for (int g = 0; g < ng; ++g) {
      float mx = -INFINITY;
      for (int e = 0; e < ne; ++e) mx = fmaxf(mx, p[c_gate + g * ne + e] + a.offset[c_gate + g * ne + e]);
      float sum = 0.f;
      for (int e = 0; e < ne; ++e) sum += expf(p[c_gate + g * ne + e] + a.offset[c_gate + g * ne + e] - mx);
      const float inv = 1.f / sum;
      for (int e = 0; e < ne; ++e)
        a.gate[b * (ng * ne) + g * ne + e] = expf(p[c_gate + g * ne + e] + a.offset[c_gate + g * ne + e] - mx) * inv;
    }
    float alpha = 1.f;
    for (int k = 0; k < nc; ++k) {
      a.alpha[b * (nc + 1) + k] = alpha;
      alpha += alpha * p[c_cross + k] + a.offset[c_cross + k];
    }
    a.alpha[b * (nc + 1) + nc] = alpha;
    for (int t = 0; t < nh; ++t) a.head[b * nh + t] = alpha * p[c_head + t] + a.offset[c_head + t];
  }
}

// gradient w.r.t. P (d_p) and w.r.t. the additive constants (d_c, to be summed over rows)
__global__ void __launch_bounds__(kThreads) rowpass_prologue_bwd_kernel(const aread_rowpass_args a) {
  const int ng = a.n_gate, ne = a.n_expert, nc = a.n_cross, nh = a.n_head;
  const int c_gate = 1, c_cross = 1 + ng * ne, c_head = c_cross + nc;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; b < a.m;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float* p = a.p + b * a.ldp;
    float* dp = a.d_p + b * a.ldp;
    float* dc = a.d_c + b * a.ldp;
    const float dl = a.d_lin ? a.d_lin[b] : 0.f;
    dp[0] = dl;
    dc[0] = dl;
    for (int g = 0; g < ng; ++g) {  // softmax backward
      float dot = 0.f;
      for (int e = 0; e < ne; ++e)
        dot += a.gate[b * (ng * ne) + g * ne + e] * (a.d_gate ? a.d_gate[b * (ng * ne) + g * ne + e] : 0.f);
      for (int e = 0; e < ne; ++e) {
        const int j = g * ne + e;
        const float v = a.gate[b * (ng * ne) + j] * ((a.d_gate ? a.d_gate[b * (ng * ne) + j] : 0.f) - dot);
        dp[c_gate + j] = v;
        dc[c_gate + j] = v;
      }
    }
    const float alpha_n = a.alpha[b * (nc + 1) + nc];
    float d_alpha = 0.f;
    for (int t = 0; t < nh; ++t) {
      const float dh = a.d_head ? a.d_head[b * nh + t] : 0.f;
      d_alpha = fmaf(dh, p[c_head + t], d_alpha);
      dp[c_head + t] = alpha_n * dh;
      dc[c_head + t] = dh;
    }
    for (int k = nc - 1; k >= 0; --k) {  // alpha_{k+1} = alpha_k + s_k, s_k = alpha_k p_k + kappa_k
      const float ds = d_alpha;
      const float alpha_k = a.alpha[b * (nc + 1) + k];
      dp[c_cross + k] = ds * alpha_k;
      dc[c_cross + k] = ds;
      d_alpha = fmaf(ds, p[c_cross + k], d_alpha);
    }
    for (int j = c_head + nh; j < a.ldp; ++j) {
      dp[j] = 0.f;
      dc[j] = 0.f;
    }
  }
}

// d_x[b, :] = sum_j d_p[b, j] * W[j, :]  (warp per 4 rows)  and the per-CTA partial of
// d_w[j, :] = sum_b d_p[b, j] * X[b, :]  (thread per column), over tiles of kTileRows rows.
constexpr int kTileRows = 32;
constexpr int kMaxColsPerThread = 6;  // E <= 1536

template <int COLS>
__global__ void __launch_bounds__(kThreads) rowdots_bwd_kernel(int64_t m, int E, int nj, int ldp,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ d_p, float* __restrict__ d_x,
                                                               float* __restrict__ dw_partial) {
  extern __shared__ float smem[];
  float* s_w = smem;                 // [nj][E]
  float* s_dp = smem + nj * E;       // [kTileRows][nj]
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, warps = blockDim.x / 32;
  for (int i = threadIdx.x; i < nj * E; i += blockDim.x) s_w[i] = w[i];
  float dw[COLS][kJChunk];
#pragma unroll
  for (int c = 0; c < COLS; ++c)
#pragma unroll
    for (int j = 0; j < kJChunk; ++j) dw[c][j] = 0.f;

  const int64_t n_tiles = (m + kTileRows - 1) / kTileRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kTileRows;
    const int rows = m - b0 < kTileRows ? static_cast<int>(m - b0) : kTileRows;
    __syncthreads();
    for (int i = threadIdx.x; i < kTileRows * nj; i += blockDim.x) {
      const int r = i / nj, j = i - r * nj;
      s_dp[i] = r < rows ? d_p[(b0 + r) * ldp + j] : 0.f;
    }
    __syncthreads();
    // ---- d_x: warp `warp` owns rows warp*4 .. warp*4+3 of the tile (8 warps x 4 rows = 32)
    if (d_x != nullptr) {
      for (int rr = warp * kRowsPerWarp; rr < kTileRows; rr += warps * kRowsPerWarp) {
        for (int e = lane; e < E; e += 32) {
          float acc[kRowsPerWarp];
#pragma unroll
          for (int r = 0; r < kRowsPerWarp; ++r) acc[r] = 0.f;
          for (int j = 0; j < nj; ++j) {
            const float wv = s_w[j * E + e];
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) acc[r] = fmaf(s_dp[(rr + r) * nj + j], wv, acc[r]);
          }
#pragma unroll
          for (int r = 0; r < kRowsPerWarp; ++r)
            if (rr + r < rows) d_x[(b0 + rr + r) * E + e] = acc[r];
        }
      }
    }
    // ---- d_w partial: thread owns columns threadIdx.x + c * blockDim.x
    for (int r = 0; r < rows; ++r) {
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int e = threadIdx.x + c * kThreads;
        if (e < E) {
          const float xv = __ldg(x + (b0 + r) * E + e);
#pragma unroll
          for (int j = 0; j < kJChunk; ++j)
            if (j < nj) dw[c][j] = fmaf(s_dp[r * nj + j], xv, dw[c][j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int e = threadIdx.x + c * kThreads;
    if (e < E) {
#pragma unroll
      for (int j = 0; j < kJChunk; ++j)
        if (j < nj) dw_partial[(static_cast<int64_t>(blockIdx.x) * nj + j) * E + e] = dw[c][j];
    }
  }
}

// d_w[j, e] = sum over CTAs (in CTA order) of the partials
__global__ void __launch_bounds__(kThreads) rowdots_bwd_reduce_kernel(int n_partial, int64_t elems,
                                                                      const float* __restrict__ partial,
                                                                      float* __restrict__ d_w) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < elems;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < n_partial; ++c) acc += partial[static_cast<int64_t>(c) * elems + i];
    d_w[i] = acc;
  }
}

int bwd_ctas(int64_t m) {
  const int64_t tiles = (m + kTileRows - 1) / kTileRows;
  return static_cast<int>(tiles < kNumSMs ? (tiles < 1 ? 1 : tiles) : kNumSMs);
}

}  // namespace
}  // namespace aread

extern "C" {

size_t aread_rowpass_workspace_bytes(int64_t m, int32_t e, int32_t n_cols) {
  return aread::align_up(static_cast<size_t>(aread::bwd_ctas(m)) * n_cols * e * 4, 256);
}

int aread_rowpass_fwd(const aread_rowpass_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "rowpass_fwd: null args");
  const aread_rowpass_args& a = *args;
  const int nj = 1 + a.n_gate * a.n_expert + a.n_cross + a.n_head;
  AREAD_REQUIRE(a.m >= 0 && a.e > 0 && nj <= a.ldp, "rowpass_fwd: bad shape (e=%d, columns=%d, ldp=%d)", a.e, nj, a.ldp);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.x && a.w && a.offset && a.p && a.lin && a.alpha, "rowpass_fwd: null pointer");
  AREAD_REQUIRE((a.gate || a.n_gate * a.n_expert == 0) && (a.head || a.n_head == 0), "rowpass_fwd: null output");
  const size_t smem = static_cast<size_t>(nj < kJChunk ? nj : kJChunk) * a.e * sizeof(float);
  AREAD_REQUIRE(smem <= 200 * 1024, "rowpass_fwd: embedding row of %d floats is too wide", a.e);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AREAD_CUDA(cudaFuncSetAttribute(rowdots_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t row_groups = (a.m + kRowsPerWarp * (kThreads / 32) - 1) / (kRowsPerWarp * (kThreads / 32));
  const unsigned grid = static_cast<unsigned>(row_groups < kNumSMs * 2 ? row_groups : kNumSMs * 2);
  AREAD_LAUNCH(rowdots_fwd_kernel, grid, kThreads, smem, stream, a.m, a.e, nj, a.x, a.w, a.p, a.ldp);
  const unsigned egrid = static_cast<unsigned>((a.m + kThreads - 1) / kThreads < kNumSMs * 8
                                                   ? (a.m + kThreads - 1) / kThreads
                                                   : kNumSMs * 8);
  AREAD_LAUNCH(rowpass_epilogue_kernel, egrid, kThreads, 0, stream, a);
  return AREAD_OK;
}

int aread_rowpass_bwd(const aread_rowpass_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "rowpass_bwd: null args");
  const aread_rowpass_args& a = *args;
  const int nj = 1 + a.n_gate * a.n_expert + a.n_cross + a.n_head;
  AREAD_REQUIRE(a.m >= 0 && a.e > 0 && nj <= a.ldp, "rowpass_bwd: bad shape");
  AREAD_REQUIRE(nj <= kJChunk, "rowpass_bwd: %d dot products exceed the supported %d", nj, kJChunk);
  AREAD_REQUIRE(a.e <= kThreads * kMaxColsPerThread, "rowpass_bwd: embedding row of %d floats is too wide", a.e);
  AREAD_REQUIRE(a.d_w && a.d_p && a.d_c && a.workspace, "rowpass_bwd: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a.m == 0) {
    AREAD_CUDA(cudaMemsetAsync(a.d_w, 0, static_cast<size_t>(nj) * a.e * 4, stream));
    return AREAD_OK;
  }
  AREAD_REQUIRE(a.x && a.w && a.p && a.alpha && (a.gate || a.n_gate * a.n_expert == 0), "rowpass_bwd: null pointer");
  const int ctas = bwd_ctas(a.m);
  AREAD_REQUIRE(a.workspace_bytes >= static_cast<size_t>(ctas) * nj * a.e * 4, "rowpass_bwd: workspace too small");
  const unsigned egrid = static_cast<unsigned>((a.m + kThreads - 1) / kThreads < kNumSMs * 8
                                                   ? (a.m + kThreads - 1) / kThreads
                                                   : kNumSMs * 8);
  AREAD_LAUNCH(rowpass_prologue_bwd_kernel, egrid, kThreads, 0, stream, a);
  const size_t smem = (static_cast<size_t>(nj) * a.e + static_cast<size_t>(kTileRows) * nj) * sizeof(float);
  AREAD_REQUIRE(smem <= 200 * 1024, "rowpass_bwd: embedding row of %d floats is too wide", a.e);
  float* partial = static_cast<float*>(a.workspace);
#define AREAD_ROWDOTS_BWD(C)                                                                                          \
  do {                                                                                                                \
    AREAD_CUDA(cudaFuncSetAttribute(rowdots_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    AREAD_LAUNCH(rowdots_bwd_kernel<C>, ctas, kThreads, smem, stream, a.m, a.e, nj, a.ldp, a.x, a.w, a.d_p, a.d_x,     \
                 partial);                                                                                            \
  } while (0)
  switch ((a.e + kThreads - 1) / kThreads) {
    case 1: AREAD_ROWDOTS_BWD(1); break;
    case 2: AREAD_ROWDOTS_BWD(2); break;
    case 3: AREAD_ROWDOTS_BWD(3); break;
    case 4: AREAD_ROWDOTS_BWD(4); break;
    case 5: AREAD_ROWDOTS_BWD(5); break;
    default: AREAD_ROWDOTS_BWD(6); break;
  }
#undef AREAD_ROWDOTS_BWD
  const int64_t elems = static_cast<int64_t>(nj) * a.e;
  AREAD_LAUNCH(rowdots_bwd_reduce_kernel, static_cast<unsigned>((elems + kThreads - 1) / kThreads), kThreads, 0, stream,
               ctas, elems, partial, a.d_w);
  return AREAD_OK;
}

}  // extern "C"
