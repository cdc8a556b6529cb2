// Mixed-domain batches: the HEI levels for rows of MANY domains in one launch, every row under the HEMP mask of its
// own domain, and per-domain means of the gate values.
//
// The reference runs the masked modes one domain at a time (model/aread.py:224-244 takes one `domain_i`; the test
// loop calls the model once per domain loader, run.py:719-727) and records unmasked gate values with a Python loop
// over domains (aread.py:187-200).  In eval mode every op of the path is row-local (BatchNorm uses the running
// statistics, dropout is off, the gates depend on the row's own domain embedding and its domain's mask), so a
// domain-sorted -- or arbitrarily mixed -- batch gives, row for row, what the per-domain calls give.
//
// aread_hei_mixed_eval: one persistent CTA per SM keeps ALL tower weights (BatchNorm folded in while they are
// staged) in shared memory and walks 32-row tiles through level 0 -> gate mix -> level 1 -> gate mix -> level 2 ->
// sigmoid heads -> mean over the row's active heads.  A (tile, tower) pair whose tower is pruned for every row of the
// tile is skipped (rows sorted by domain make that the common case); inside a tile a pruned tower's output is never
// read (its mixing weight is zero, like the zeros the reference substitutes, aread.py:299-300).
//
// aread_domain_mean: mean over the rows of each domain of a [m, C] matrix, rows added in batch order (deterministic).
#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 32;          // rows per tile
constexpr int kMaxLevel = AREAD_MIXED_MAX_LEVEL;
constexpr int kMaxLayer = AREAD_MIXED_MAX_LAYER;

struct Plan {                      // shared-memory offsets (floats), computed on the host
  int w_off[kMaxLevel][kMaxLayer];   // folded weights, [tower][k][n] (transposed for the inner loop)
  int b_off[kMaxLevel][kMaxLayer];   // folded biases  [tower][n]
  int tail_off;                      // head weights [n_last][w_last]
  int buf_a, buf_b;                  // activation ping-pong, [kRows][ld]
  int gate_off;                      // mixing weights [kRows][max n_l * n_{l-1}]
  int ld;                            // activation row stride (== 1 mod 8: conflict-free row-quad reads)
  int total;                         // floats
};

// out[r, t, n] = relu(sum_k in[r, t, k] * W'[t][k][n] + b'[t][n]) for the towers that run somewhere in the tile
__device__ __forceinline__ void tile_layer(const float* __restrict__ sIn, float* __restrict__ sOut, const float* __restrict__ sW,
                                           const float* __restrict__ sB, int n_towers, int K, int N, int ld,
                                           uint32_t tile_active) {
  const int nq = N / 4;
  const int items = n_towers * nq * (kRows / 4);
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int rq = it % (kRows / 4);
    const int cq = (it / (kRows / 4)) % nq;
    const int t = it / ((kRows / 4) * nq);
    if (!((tile_active >> t) & 1u)) continue;
    const float* in = sIn + (rq * 4) * ld + t * K;
    const float* w = sW + t * K * N + cq * 4;
    float acc[4][4];
    const float4 b4 = *reinterpret_cast<const float4*>(sB + t * N + cq * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r) { acc[r][0] = b4.x; acc[r][1] = b4.y; acc[r][2] = b4.z; acc[r][3] = b4.w; }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + k * N);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float v = in[r * ld + k];
        acc[r][0] = fmaf(v, w4.x, acc[r][0]);
        acc[r][1] = fmaf(v, w4.y, acc[r][1]);
        acc[r][2] = fmaf(v, w4.z, acc[r][2]);
        acc[r][3] = fmaf(v, w4.w, acc[r][3]);
      }
    }
    float* out = sOut + (rq * 4) * ld + t * N + cq * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) out[r * ld + c] = fmaxf(acc[r][c], 0.f);
  }
}

__global__ void __launch_bounds__(kThreads, 1) hei_mixed_eval_kernel(const aread_hei_mixed_args a, const Plan pl) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int sDom[kRows];
  __shared__ uint32_t sTileAct[kMaxLevel];
  const int n_level = a.n_level, n_layer = a.n_layer;

  // ---- stage every tower's weights once, BatchNorm (running statistics) folded in:
  //      W'[t][k][n] = W[t][n][k] * scale[t][n],  b'[t][n] = b[t][n] * scale[t][n] + shift[t][n]
  for (int l = 0; l < n_level; ++l) {
    int K = l == 0 ? a.width_in : a.dims[l - 1][n_layer - 1];
    for (int j = 0; j < n_layer; ++j) {
      const int N = a.dims[l][j], T = a.n_tower[l];
      float* sW = smem + pl.w_off[l][j];
      float* sB = smem + pl.b_off[l][j];
      for (int idx = threadIdx.x; idx < T * N * K; idx += kThreads) {
        const int t = idx / (N * K), rem = idx - t * N * K;
        const int n = rem / K, k = rem - n * K;
        float scale = 1.f;
        if (!a.bn_skip) scale = __ldg(a.gamma[l][j] + t * N + n) * rsqrtf(__ldg(a.running_var[l][j] + t * N + n) + a.eps);
        sW[t * K * N + k * N + n] = __ldg(a.weight[l][j] + idx) * scale;
      }
      for (int idx = threadIdx.x; idx < T * N; idx += kThreads) {
        float scale = 1.f, shift = 0.f;
        if (!a.bn_skip) {
          scale = __ldg(a.gamma[l][j] + idx) * rsqrtf(__ldg(a.running_var[l][j] + idx) + a.eps);
          shift = __ldg(a.beta[l][j] + idx) - __ldg(a.running_mean[l][j] + idx) * scale;
        }
        sB[idx] = __ldg(a.bias[l][j] + idx) * scale + shift;
      }
      K = N;
    }
  }
  const int n_last = a.n_tower[n_level - 1], w_last = a.dims[n_level - 1][n_layer - 1];
  for (int idx = threadIdx.x; idx < n_last * w_last; idx += kThreads) smem[pl.tail_off + idx] = __ldg(a.w_tail + idx);
  __syncthreads();

  float* bufA = smem + pl.buf_a;
  float* bufB = smem + pl.buf_b;
  float* sGate = smem + pl.gate_off;
  const int ld = pl.ld;
  const int64_t n_tiles = (a.m + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kRows;
    const int rows = a.m - b0 < kRows ? static_cast<int>(a.m - b0) : kRows;
    __syncthreads();
    if (threadIdx.x < kRows) {
      int d = threadIdx.x < rows ? __ldg(a.domain + (b0 + threadIdx.x) * a.domain_stride) : -1;
      if (d >= a.n_domain) d = -1;                       // out-of-range ids run no tower
      sDom[threadIdx.x] = d;
    }
    if (threadIdx.x < kMaxLevel) sTileAct[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x < kRows && sDom[threadIdx.x] >= 0)
      for (int l = 0; l < n_level; ++l) atomicOr(&sTileAct[l], __ldg(a.active + sDom[threadIdx.x] * n_level + l));
    // level-0 inputs: [rows, n0 * width_in]
    {
      const int w0 = a.n_tower[0] * a.width_in;
      for (int idx = threadIdx.x; idx < kRows * w0; idx += kThreads) {
        const int r = idx / w0, c = idx - r * w0;
        bufA[r * ld + c] = r < rows ? __ldg(a.t0 + (b0 + r) * w0 + c) : 0.f;
      }
    }
    __syncthreads();
    float* cur = bufA;
    float* nxt = bufB;
    int edge_base = 0;
    for (int l = 0; l < n_level; ++l) {
      const int T = a.n_tower[l];
      if (l > 0) {
        // ---- gate mix: softmax over the previous level's towers, masked by the row's domain, renormalised
        //      (aread.py:282-288); pruned previous towers contribute their zeros
        const int P = a.n_tower[l - 1], Wp = a.dims[l - 1][n_layer - 1];
        for (int it = threadIdx.x; it < kRows * T; it += kThreads) {
          const int r = it / T, t = it - r * T;
          float* g = sGate + it * P;
          const int d = sDom[r];
          if (d < 0) {
            for (int j = 0; j < P; ++j) g[j] = 0.f;
            continue;
          }
          const uint32_t em = __ldg(a.edges + static_cast<int64_t>(d) * a.edge_words + edge_base + t);
          const uint32_t act_prev = __ldg(a.active + d * n_level + l - 1);
          const float* lg = a.logits[l] + ((b0 + r) * T + t) * P;
          float mx = -INFINITY;
          for (int j = 0; j < P; ++j) { g[j] = __ldg(lg + j); mx = fmaxf(mx, g[j]); }
          float sum = 0.f;
          for (int j = 0; j < P; ++j) { g[j] = expf(g[j] - mx); sum += g[j]; }
          const float inv = 1.f / sum;
          float tot = 0.f;
          for (int j = 0; j < P; ++j) { g[j] = g[j] * inv * (((em >> j) & 1u) ? 1.f : 0.f); tot += g[j]; }
          const float denom = tot + 1e-8f;
          for (int j = 0; j < P; ++j) g[j] = ((act_prev >> j) & 1u) ? g[j] / denom : 0.f;
        }
        __syncthreads();
        const int OW = T * Wp;
        for (int idx = threadIdx.x; idx < kRows * OW; idx += kThreads) {
          const int r = idx / OW, rem = idx - r * OW;
          const int t = rem / Wp, c = rem - t * Wp;
          const float* g = sGate + (r * T + t) * P;
          float acc = 0.f;
          for (int j = 0; j < P; ++j)
            if (g[j] != 0.f) acc = fmaf(g[j], cur[r * ld + j * Wp + c], acc);   // pruned towers were never written
          nxt[r * ld + rem] = acc;
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
        edge_base += T;
      }
      int K = l == 0 ? a.width_in : a.dims[l - 1][n_layer - 1];
      const uint32_t tile_act = sTileAct[l];
      for (int j = 0; j < n_layer; ++j) {
        const int N = a.dims[l][j];
        tile_layer(cur, nxt, smem + pl.w_off[l][j], smem + pl.b_off[l][j], T, K, N, ld, tile_act);
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
        K = N;
      }
    }
    // ---- heads: p_t = sigmoid(head_cross[b, t] + <u_t, w_tail[t]> + lin[b]) for the row's active towers; y = mean
    if (threadIdx.x < rows) {
      const int r = threadIdx.x;
      const int d = sDom[r];
      const int64_t b = b0 + r;
      const uint32_t act = d >= 0 ? __ldg(a.active + d * n_level + n_level - 1) : 0u;
      const float lin = __ldg(a.lin + b);
      float sum = 0.f;
      int cnt = 0;
      for (int t = 0; t < n_last; ++t) {
        float p = 0.f;
        if ((act >> t) & 1u) {
          float z = __ldg(a.head_cross + b * n_last + t) + lin;
          const float* u = cur + r * ld + t * w_last;
          const float* w = smem + pl.tail_off + t * w_last;
          for (int c = 0; c < w_last; ++c) z = fmaf(u[c], w[c], z);
          p = 1.f / (1.f + expf(-z));
          sum += p;
          ++cnt;
        }
        if (a.y_stack != nullptr) a.y_stack[static_cast<int64_t>(t) * a.m + b] = p;
      }
      a.y[b] = cnt > 0 ? sum / static_cast<float>(cnt) : 0.f;
    }
  }
}

int make_plan(const aread_hei_mixed_args& a, Plan* pl) {
  int off = 0, max_w = a.n_tower[0] * a.width_in, max_gate = 1;
  for (int l = 0; l < a.n_level; ++l) {
    int K = l == 0 ? a.width_in : a.dims[l - 1][a.n_layer - 1];
    if (l > 0) {
      if (a.n_tower[l] * K > max_w) max_w = a.n_tower[l] * K;
      if (a.n_tower[l] * a.n_tower[l - 1] > max_gate) max_gate = a.n_tower[l] * a.n_tower[l - 1];
    }
    for (int j = 0; j < a.n_layer; ++j) {
      const int N = a.dims[l][j];
      if (N % 4 != 0 || N <= 0) return -1;
      pl->w_off[l][j] = off;
      off += a.n_tower[l] * K * N;
      if (a.n_tower[l] * N > max_w) max_w = a.n_tower[l] * N;
      K = N;
    }
  }
  for (int l = 0; l < a.n_level; ++l)
    for (int j = 0; j < a.n_layer; ++j) {
      pl->b_off[l][j] = off;
      off += a.n_tower[l] * a.dims[l][j];
    }
  off = (off + 3) & ~3;
  pl->tail_off = off;
  off += a.n_tower[a.n_level - 1] * a.dims[a.n_level - 1][a.n_layer - 1];
  off = (off + 3) & ~3;
  int ld = max_w;
  while (ld % 8 != 1) ++ld;
  pl->ld = ld;
  pl->buf_a = off;
  off += kRows * ld;
  pl->buf_b = off;
  off += kRows * ld;
  pl->gate_off = off;
  off += kRows * max_gate;
  pl->total = off;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// mean over the rows of each domain: one CTA per domain scans the id column with warp ballots and adds the matching
// rows in batch order; thread = column
__global__ void __launch_bounds__(kThreads) domain_mean_kernel(const aread_domain_mean_args a) {
  const int d = blockIdx.x;
  __shared__ int sRows[kThreads];
  __shared__ int sCount;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  constexpr int kWarps = kThreads / 32;
  __shared__ int sWarpCount[kWarps];
  int total = 0;
  float acc[AREAD_DOMAIN_MEAN_MAX_COLS / kThreads];
#pragma unroll
  for (int c = 0; c < AREAD_DOMAIN_MEAN_MAX_COLS / kThreads; ++c) acc[c] = 0.f;
  for (int64_t base = 0; base < a.m; base += kThreads) {
    const int64_t b = base + threadIdx.x;
    const bool hit = b < a.m && __ldg(a.domain + b * a.domain_stride) == d;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) sWarpCount[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < kWarps; ++w) { if (w < warp) before += sWarpCount[w]; all += sWarpCount[w]; }
    if (hit) sRows[before + __popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();
    for (int i = 0; i < all; ++i) {               // matching rows of this block of ids, in batch order
      const float* row = a.values + (base + sRows[i]) * a.ld;
#pragma unroll
      for (int c = 0; c < AREAD_DOMAIN_MEAN_MAX_COLS / kThreads; ++c) {
        const int col = c * kThreads + threadIdx.x;
        if (col < a.width) acc[c] += __ldg(row + col);
      }
    }
    total += all;
    __syncthreads();
  }
  if (threadIdx.x == 0 && a.count != nullptr) a.count[d] = total;
#pragma unroll
  for (int c = 0; c < AREAD_DOMAIN_MEAN_MAX_COLS / kThreads; ++c) {
    const int col = c * kThreads + threadIdx.x;
    if (col < a.width) a.mean[static_cast<int64_t>(d) * a.width + col] = total > 0 ? acc[c] / static_cast<float>(total) : 0.f;
  }
  (void)sCount;
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_hei_mixed_eval_supported(const aread_hei_mixed_args* args) {
  using namespace aread;
  if (args == nullptr) return 0;
  const aread_hei_mixed_args& a = *args;
  if (a.n_level < 1 || a.n_level > kMaxLevel || a.n_layer < 1 || a.n_layer > kMaxLayer) return 0;
  for (int l = 0; l < a.n_level; ++l)
    if (a.n_tower[l] < 1 || a.n_tower[l] > 32) return 0;
  if (a.width_in % 4 != 0) return 0;
  Plan pl;
  if (make_plan(a, &pl) != 0) return 0;
  return static_cast<size_t>(pl.total) * 4 <= 220 * 1024 ? 1 : 0;
}

int aread_hei_mixed_eval(const aread_hei_mixed_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "hei_mixed_eval: null args");
  const aread_hei_mixed_args& a = *args;
  AREAD_REQUIRE(aread_hei_mixed_eval_supported(args), "hei_mixed_eval: tower configuration does not fit one SM's shared memory");
  if (a.m <= 0) return AREAD_OK;
  AREAD_REQUIRE(a.domain && a.active && a.t0 && a.head_cross && a.lin && a.w_tail && a.y && a.n_domain > 0,
                "hei_mixed_eval: null pointer");
  AREAD_REQUIRE(a.n_level == 1 || a.edges != nullptr, "hei_mixed_eval: null edges");
  for (int l = 0; l < a.n_level; ++l) {
    AREAD_REQUIRE(l == 0 || a.logits[l] != nullptr, "hei_mixed_eval: null logits of level %d", l);
    for (int j = 0; j < a.n_layer; ++j)
      AREAD_REQUIRE(a.weight[l][j] && a.bias[l][j] &&
                        (a.bn_skip || (a.gamma[l][j] && a.beta[l][j] && a.running_mean[l][j] && a.running_var[l][j])),
                    "hei_mixed_eval: null parameter of level %d layer %d", l, j);
  }
  Plan pl;
  make_plan(a, &pl);
  const size_t smem = static_cast<size_t>(pl.total) * 4;
  static uint64_t configured = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(configured & (uint64_t{1} << (dev & 63)))) {
    AREAD_CUDA(cudaFuncSetAttribute(hei_mixed_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured |= uint64_t{1} << (dev & 63);
  }
  const int64_t n_tiles = (a.m + kRows - 1) / kRows;
  AREAD_LAUNCH(hei_mixed_eval_kernel, static_cast<unsigned>(n_tiles < kNumSMs ? n_tiles : kNumSMs), kThreads, smem,
               static_cast<cudaStream_t>(stream_), a, pl);
  return AREAD_OK;
}

int aread_domain_mean(const aread_domain_mean_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "domain_mean: null args");
  const aread_domain_mean_args& a = *args;
  AREAD_REQUIRE(a.n_domain > 0 && a.width > 0 && a.width <= AREAD_DOMAIN_MEAN_MAX_COLS, "domain_mean: width %d not in [1, %d]",
                a.width, AREAD_DOMAIN_MEAN_MAX_COLS);
  AREAD_REQUIRE(a.m >= 0 && a.values && a.domain && a.mean, "domain_mean: null pointer");
  AREAD_LAUNCH(domain_mean_kernel, static_cast<unsigned>(a.n_domain), kThreads, 0, static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

}  // extern "C"
