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
constexpr int kTile = 64;    // rows per tile
constexpr int kChunk = 64;   // embedding elements per shared-memory chunk
constexpr int kJ = 32;       // dot products handled per launch
constexpr int kXs = kChunk + 1;

// stage X[b0 .. b0+64, e0 .. e0+64) (zero padded) into sX[r][e] with a conflict-free row stride
__device__ __forceinline__ void load_x_chunk(float* sX, const float* __restrict__ x, int64_t m, int E, int64_t b0,
                                             int e0) {
  for (int idx = threadIdx.x; idx < kTile * kChunk; idx += kThreads) {
    const int r = idx / kChunk, e = idx - r * kChunk;
    sX[r * kXs + e] = (b0 + r < m && e0 + e < E) ? __ldg(x + (b0 + r) * E + e0 + e) : 0.f;
  }
}

// P[b, j] = sum_e X[b, e] * W[j, e], j < nj <= 32.  CTA = 64-row tiles; thread = (row, 8 dot products).
// The next 64-element chunk of X is fetched into registers while the current one is multiplied (the loads are the
// long pole: one pass over X), and W is staged with the lanes running along j so that the transposed
// shared-memory stores do not collide.
__global__ void __launch_bounds__(kThreads, 3) rowdots_fwd_kernel(int64_t m, int E, int nj, const float* __restrict__ x,
                                                               const float* __restrict__ w, float* __restrict__ p,
                                                               int ldp) {
  __shared__ float sX[kTile * kXs];
  __shared__ __align__(16) float sW[kChunk * kJ];  // [e][j]
  const int r = threadIdx.x / 4, jq = threadIdx.x % 4;
  const int lr = threadIdx.x / kChunk, le = threadIdx.x % kChunk;      // loader role: rows lr, lr + 4, ...; column le
  const int wj = threadIdx.x % kJ, we = threadIdx.x / kJ;             // W loader: column wj, rows we, we + 8, ...
  const int64_t n_tiles = (m + kTile - 1) / kTile;
  const int n_chunks = (E + kChunk - 1) / kChunk;
  float pre[kTile / 4];
  auto fetch = [&](int64_t b0, int e0) {
#pragma unroll
    for (int k = 0; k < kTile / 4; ++k) {
      const int64_t row = b0 + lr + 4 * k;
      pre[k] = (row < m && e0 + le < E) ? __ldg(x + row * E + e0 + le) : 0.f;
    }
  };
  int64_t tile = blockIdx.x;
  if (tile < n_tiles) fetch(tile * kTile, 0);
  for (; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kTile;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      const int e0 = c * kChunk;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTile / 4; ++k) sX[(lr + 4 * k) * kXs + le] = pre[k];
#pragma unroll
      for (int k = 0; k < kChunk / 8; ++k) {
        const int e = we + 8 * k;
        sW[e * kJ + wj] = (wj < nj && e0 + e < E) ? __ldg(w + static_cast<int64_t>(wj) * E + e0 + e) : 0.f;
      }
      __syncthreads();
      // prefetch the next chunk (of this tile, or the first chunk of the CTA's next tile)
      if (c + 1 < n_chunks) fetch(b0, e0 + kChunk);
      else if (tile + gridDim.x < n_tiles) fetch((tile + gridDim.x) * kTile, 0);
#pragma unroll 8
      for (int e = 0; e < kChunk; ++e) {
        const float xv = sX[r * kXs + e];
        const float4 w0 = *reinterpret_cast<const float4*>(sW + e * kJ + jq * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sW + e * kJ + jq * 8 + 4);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
        acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
    if (b0 + r < m) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (jq * 8 + j < nj) p[(b0 + r) * ldp + jq * 8 + j] = acc[j];
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

// gradient w.r.t. P (d_p) and w.r.t. the additive constants (d_c, to be summed over rows).  A CTA owns kPrologueRows
// rows: one thread per row works the scalar recurrences out into shared memory, then all threads write d_p, d_c and
// the split bf16 copy with consecutive threads on consecutive columns (whole 128-byte lines per row).
constexpr int kPrologueRows = 32;

__global__ void __launch_bounds__(kThreads) rowpass_prologue_bwd_kernel(const aread_rowpass_args a) {
  extern __shared__ float s_dp[];                       // [kPrologueRows][ldp + 1]: d_p, then reused for d_c deltas
  const int ng = a.n_gate, ne = a.n_expert, nc = a.n_cross, nh = a.n_head;
  const int c_gate = 1, c_cross = 1 + ng * ne, c_head = c_cross + nc;
  const int n_own = c_head + nh, n_all = n_own + a.n_extra;
  const int ld = a.ldp, lds = a.ldp + 1;
  float* s_dc = s_dp + kPrologueRows * lds;             // [kPrologueRows][ldp + 1]
  float col_sum = 0.f;                                  // of d_c column threadIdx.x over this CTA's tiles, in row order
  const int64_t n_tiles = (a.m + kPrologueRows - 1) / kPrologueRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kPrologueRows;
    const int rows = a.m - b0 < kPrologueRows ? static_cast<int>(a.m - b0) : kPrologueRows;
    __syncthreads();
    // riding dot products: their d_p was written by the owner (aread_gate_mix); bring it in, whole rows at a time
    for (int idx = threadIdx.x; idx < rows * a.n_extra; idx += kThreads) {
      const int r = idx / a.n_extra, j = n_own + (idx - r * a.n_extra);
      const float v = a.d_p[(b0 + r) * ld + j];
      s_dp[r * lds + j] = v;
      s_dc[r * lds + j] = v;
    }
    if (threadIdx.x < rows) {
      const int r = threadIdx.x;
      const int64_t b = b0 + r;
      const float* p = a.p + b * ld;
      float* dp = s_dp + r * lds;
      float* dc = s_dc + r * lds;
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
      for (int j = n_all; j < ld; ++j) {
        dp[j] = 0.f;
        dc[j] = 0.f;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * ld; idx += kThreads) {
      const int r = idx / ld, j = idx - r * ld;
      a.d_p[(b0 + r) * ld + j] = s_dp[r * lds + j];
      if (a.d_c != nullptr) a.d_c[(b0 + r) * ld + j] = s_dc[r * lds + j];
    }
    if (a.d_c_sum != nullptr && threadIdx.x < ld) {
      for (int r = 0; r < rows; ++r) col_sum += s_dc[r * lds + threadIdx.x];
    }
    if (a.dp16 != nullptr) {   // split bf16 operands of the tensor-core products: [hi | hi | lo], W columns each
      const int W = a.dp16_width > 0 ? a.dp16_width : 32;
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.dp16);
      for (int idx = threadIdx.x; idx < rows * W; idx += kThreads) {
        const int r = idx / W, j = idx - r * W;
        const float v = j < n_all ? s_dp[r * lds + j] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16* row = o + (b0 + r) * a.ld16;
        row[j] = hi;
        row[W + j] = hi;
        row[2 * W + j] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    }
  }
  if (a.d_c_sum != nullptr && threadIdx.x < ld) a.d_c_partial[static_cast<int64_t>(blockIdx.x) * ld + threadIdx.x] = col_sum;
}

// d_x[b, e] (+)= sum_j d_p[b, j] * W[j, e]  and the per-CTA partial of  d_w[j, e] = sum_b d_p[b, j] * X[b, e].
// CTA = 64-row tiles, the embedding row is walked in 64-element chunks staged in shared memory.
//   d_x phase: thread = (row, 16 consecutive e);   d_w phase: thread = (e, 8 of the 32 dot products),
//   accumulating over all tiles of the CTA in registers (MAX_CHUNKS x 8).
template <int MAX_CHUNKS>
__global__ void __launch_bounds__(kThreads, MAX_CHUNKS <= 5 ? 2 : 1) rowdots_bwd_kernel(int64_t m, int E, int nj, int ldp,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ d_p, float* __restrict__ d_x,
                                                               int accumulate_dx, float* __restrict__ dw_partial) {
  __shared__ float sX[kTile * kXs];
  __shared__ __align__(16) float sW[kJ * kChunk];   // [j][e]
  __shared__ __align__(16) float sDp[kTile * kJ];   // [r][j]
  const int r = threadIdx.x / 4, eq = threadIdx.x % 4;   // d_x phase
  const int ec = threadIdx.x % kChunk, jg = threadIdx.x / kChunk;  // d_w phase: 4 groups of 8 j
  float dw[MAX_CHUNKS][8];
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[c][j] = 0.f;

  const int64_t n_tiles = (m + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kTile;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kTile * kJ; idx += kThreads) {
      const int rr = idx / kJ, j = idx - rr * kJ;
      sDp[idx] = (b0 + rr < m && j < nj) ? __ldg(d_p + (b0 + rr) * ldp + j) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c) {
      const int e0 = c * kChunk;
      if (e0 >= E) break;
      __syncthreads();
      load_x_chunk(sX, x, m, E, b0, e0);
      for (int idx = threadIdx.x; idx < kJ * kChunk; idx += kThreads) {
        const int j = idx / kChunk, e = idx - j * kChunk;
        sW[idx] = (j < nj && e0 + e < E) ? __ldg(w + static_cast<int64_t>(j) * E + e0 + e) : 0.f;
      }
      __syncthreads();
      if (d_x != nullptr) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int j = 0; j < nj; ++j) {
          const float dp = sDp[r * kJ + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(sW + j * kChunk + eq * 16 + q * 4);
            acc[q * 4 + 0] = fmaf(dp, wv.x, acc[q * 4 + 0]);
            acc[q * 4 + 1] = fmaf(dp, wv.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(dp, wv.z, acc[q * 4 + 2]);
            acc[q * 4 + 3] = fmaf(dp, wv.w, acc[q * 4 + 3]);
          }
        }
        if (b0 + r < m) {
          float* dst = d_x + (b0 + r) * E + e0 + eq * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (e0 + eq * 16 + i < E) dst[i] = accumulate_dx ? dst[i] + acc[i] : acc[i];
        }
      }
      const int rows = m - b0 < kTile ? static_cast<int>(m - b0) : kTile;
      for (int rr = 0; rr < rows; ++rr) {
        const float xv = sX[rr * kXs + ec];
        const float4 d0 = *reinterpret_cast<const float4*>(sDp + rr * kJ + jg * 8);
        const float4 d1 = *reinterpret_cast<const float4*>(sDp + rr * kJ + jg * 8 + 4);
        dw[c][0] = fmaf(d0.x, xv, dw[c][0]); dw[c][1] = fmaf(d0.y, xv, dw[c][1]);
        dw[c][2] = fmaf(d0.z, xv, dw[c][2]); dw[c][3] = fmaf(d0.w, xv, dw[c][3]);
        dw[c][4] = fmaf(d1.x, xv, dw[c][4]); dw[c][5] = fmaf(d1.y, xv, dw[c][5]);
        dw[c][6] = fmaf(d1.z, xv, dw[c][6]); dw[c][7] = fmaf(d1.w, xv, dw[c][7]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c) {
    const int e = c * kChunk + ec;
    if (e < E) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (jg * 8 + j < nj) dw_partial[(static_cast<int64_t>(blockIdx.x) * nj + jg * 8 + j) * E + e] = dw[c][j];
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

// out[col] = sum over the CTA partials of column col: 32 columns x 8 interleaved groups of partials per block, four
// independent sums per thread, then groups in order (fixed order; ~n_partial / 32 dependent steps instead of n_partial)
__global__ void __launch_bounds__(kThreads) column_partials_reduce_kernel(int n_partial, int ld,
                                                                          const float* __restrict__ partial,
                                                                          float* __restrict__ out) {
  __shared__ float s_part[8][32];
  const int cx = threadIdx.x % 32, py = threadIdx.x / 32;
  const int col = blockIdx.x * 32 + cx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < ld) {
    int i = py;
    for (; i + 24 < n_partial; i += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += partial[static_cast<int64_t>(i + 8 * u) * ld + col];
    }
    for (int u = 0; i < n_partial; i += 8, ++u) acc[u & 3] += partial[static_cast<int64_t>(i) * ld + col];
  }
  s_part[py][cx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (py == 0 && col < ld) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += s_part[y][cx];
    out[col] = v;
  }
}

unsigned prologue_grid(int64_t m) {
  const int64_t tiles = (m + kPrologueRows - 1) / kPrologueRows;
  return static_cast<unsigned>(tiles < kNumSMs * 4 ? (tiles < 1 ? 1 : tiles) : kNumSMs * 4);
}
size_t prologue_smem(int ldp) { return sizeof(float) * 2 * kPrologueRows * (ldp + 1); }

int launch_prologue(const aread_rowpass_args& a, cudaStream_t stream) {
  AREAD_REQUIRE(a.d_c != nullptr || a.d_c_sum != nullptr, "rowpass_bwd: neither d_c nor d_c_sum given");
  AREAD_REQUIRE(a.d_c_sum == nullptr || (a.d_c_partial != nullptr && a.ldp <= kThreads),
                "rowpass_bwd: d_c_sum needs d_c_partial and at most %d columns", kThreads);
  const unsigned grid = prologue_grid(a.m);
  AREAD_LAUNCH(rowpass_prologue_bwd_kernel, grid, kThreads, prologue_smem(a.ldp), stream, a);
  if (a.d_c_sum != nullptr)
    AREAD_LAUNCH(column_partials_reduce_kernel, ceil_div(a.ldp, 32), kThreads, 0, stream, static_cast<int>(grid), a.ldp,
                 a.d_c_partial, a.d_c_sum);
  return AREAD_OK;
}

int bwd_ctas(int64_t m) {
  const int64_t tiles = (m + kTile - 1) / kTile;
  const int64_t cap = kNumSMs * 2;
  return static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
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
  AREAD_REQUIRE(a.m >= 0 && a.e > 0 && nj + a.n_extra <= a.ldp && a.n_extra >= 0,
                "rowpass_fwd: bad shape (e=%d, columns=%d + %d, ldp=%d)", a.e, nj, a.n_extra, a.ldp);
  AREAD_REQUIRE(a.n_extra == 0 || a.x == nullptr, "rowpass_fwd: riding dot products need the tensor-core variant (x == NULL)");
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE((a.x == nullptr || a.w) && a.offset && a.p && a.lin && a.alpha, "rowpass_fwd: null pointer");
  AREAD_REQUIRE((a.gate || a.n_gate * a.n_expert == 0) && (a.head || a.n_head == 0), "rowpass_fwd: null output");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a.x != nullptr) {
    const int64_t tiles = (a.m + kTile - 1) / kTile;
    const unsigned grid = static_cast<unsigned>(tiles < kNumSMs * 4 ? tiles : kNumSMs * 4);
    for (int j0 = 0; j0 < nj; j0 += kJ)  // 32 dot products per launch
      AREAD_LAUNCH(rowdots_fwd_kernel, grid, kThreads, 0, stream, a.m, a.e, nj - j0 < kJ ? nj - j0 : kJ, a.x,
                   a.w + static_cast<int64_t>(j0) * a.e, a.p + j0, a.ldp);
  }
  const unsigned egrid = static_cast<unsigned>((a.m + kThreads - 1) / kThreads < kNumSMs * 8
                                                   ? (a.m + kThreads - 1) / kThreads
                                                   : kNumSMs * 8);
  AREAD_LAUNCH(rowpass_epilogue_kernel, egrid, kThreads, 0, stream, a);
  return AREAD_OK;
}

int32_t aread_rowpass_prologue_ctas(int64_t m) { return static_cast<int32_t>(aread::prologue_grid(m)); }

int aread_rowpass_bwd(const aread_rowpass_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "rowpass_bwd: null args");
  const aread_rowpass_args& a = *args;
  const int nj = 1 + a.n_gate * a.n_expert + a.n_cross + a.n_head;
  AREAD_REQUIRE(a.m >= 0 && a.e > 0 && a.n_extra >= 0 && nj + a.n_extra <= a.ldp, "rowpass_bwd: bad shape");
  AREAD_REQUIRE(prologue_smem(a.ldp) <= 48 * 1024, "rowpass_bwd: %d dot products per row exceed the prologue's shared memory",
                a.ldp);
  AREAD_REQUIRE(a.n_extra == 0 || a.x == nullptr, "rowpass_bwd: riding dot products need the tensor-core variant (x == NULL)");
  AREAD_REQUIRE(a.e <= 16 * kChunk, "rowpass_bwd: embedding row of %d floats is too wide (max %d)", a.e, 16 * kChunk);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned pgrid = static_cast<unsigned>((a.m + kThreads - 1) / kThreads < kNumSMs * 8
                                                   ? (a.m + kThreads - 1) / kThreads
                                                   : kNumSMs * 8);
  if (a.x == nullptr) {  // per-row prologue only: the products run on the tensor cores
    AREAD_REQUIRE(a.d_p, "rowpass_bwd: null pointer");
    {
      const int w16 = a.dp16_width > 0 ? a.dp16_width : 32;
      AREAD_REQUIRE(a.dp16 == nullptr || (w16 % 32 == 0 && nj + a.n_extra <= w16 && a.ld16 >= 3 * w16),
                    "rowpass_bwd: dp16 thirds of %d columns cannot hold %d dot products", w16, nj + a.n_extra);
    }
    if (a.m == 0) return AREAD_OK;
    AREAD_REQUIRE(a.p && a.alpha && (a.gate || a.n_gate * a.n_expert == 0), "rowpass_bwd: null pointer");
    return launch_prologue(a, stream);
  }
  AREAD_REQUIRE(a.d_w && a.d_p && a.workspace, "rowpass_bwd: null pointer");
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
  if (int rc = launch_prologue(a, stream)) return rc;
  float* partial = static_cast<float*>(a.workspace);
  const int n_chunks = (a.e + kChunk - 1) / kChunk;
  for (int j0 = 0; j0 < nj; j0 += kJ) {
    const int njb = nj - j0 < kJ ? nj - j0 : kJ;
    const float* w = a.w + static_cast<int64_t>(j0) * a.e;
    float* part = partial + static_cast<int64_t>(ctas) * j0 * a.e;
    const int acc_dx = j0 > 0 ? 1 : 0;
#define AREAD_ROWDOTS_BWD(C) \
  AREAD_LAUNCH(rowdots_bwd_kernel<C>, ctas, kThreads, 0, stream, a.m, a.e, njb, a.ldp, a.x, w, a.d_p + j0, a.d_x, \
               acc_dx, part)
    if (n_chunks <= 2) AREAD_ROWDOTS_BWD(2);
    else if (n_chunks <= 5) AREAD_ROWDOTS_BWD(5);
    else if (n_chunks <= 8) AREAD_ROWDOTS_BWD(8);
    else if (n_chunks <= 12) AREAD_ROWDOTS_BWD(12);
    else AREAD_ROWDOTS_BWD(16);
#undef AREAD_ROWDOTS_BWD
    const int64_t elems = static_cast<int64_t>(njb) * a.e;
    AREAD_LAUNCH(rowdots_bwd_reduce_kernel, static_cast<unsigned>((elems + kThreads - 1) / kThreads), kThreads, 0,
                 stream, ctas, elems, part, a.d_w + static_cast<int64_t>(j0) * a.e);
  }
  return AREAD_OK;
}

}  // extern "C"
