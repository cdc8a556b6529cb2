// Grouped small Linear layers of the HEI towers in fp32 on the CUDA cores: forward, data gradient
// and weight gradient.  A tower layer is 16..64 wide (model/aread.py:106-117, tower_dims
// ((64,32),(32,16),(16,8))), far below a tensor-core tile, and the towers are ~3 % of the FLOPs; what
// matters is that ALL towers of a level run in one launch and that each activation row is read once.
//
//   out[b, g, j] = sum_i in[b, g, i] * M_g[i, j] (+ bias[g, j])
//
// forward:        in = tower input [m, G, K],  M_g = W_g^T (W_g is nn.Linear's [N, K]),  out = z [m, G, N]
// data gradient:  in = dz [m, G, N],           M_g = W_g,                                out = d_in [m, G, K]
// Towers pruned by the HEMP mask are simply absent: the caller passes the compact list of active towers.
#include "common.cuh"
#include "tensor_map.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------- forward / dgrad
// CTA = (row tile, group).  sM[i][j] holds the group's matrix, sIn the row tile; every thread owns a
// 4 x 4 (rows x columns) block of the output tile.
__global__ void __launch_bounds__(kThreads) tower_linear_kernel(const aread_tower_linear_args a, int tx_n, int ty_n) {
  extern __shared__ float smem[];
  const int I = a.in_width, J = a.out_width;
  const int Jp = (J + 3) & ~3;
  float* sM = smem;                    // [I][Jp]
  float* sIn = smem + I * Jp;          // [tile_rows][I + 1]
  const int tile_rows = ty_n * 4;
  const int g = blockIdx.y;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * tile_rows;
  const float* __restrict__ w = a.weight + static_cast<int64_t>(g) * I * J;

  for (int idx = threadIdx.x; idx < I * Jp; idx += kThreads) {
    const int i = idx / Jp, j = idx - i * Jp;
    float v = 0.f;
    if (j < J) v = a.weight_is_out_by_in ? __ldg(w + static_cast<int64_t>(j) * I + i) : __ldg(w + static_cast<int64_t>(i) * J + j);
    sM[idx] = v;
  }
  for (int idx = threadIdx.x; idx < tile_rows * I; idx += kThreads) {
    const int r = idx / I, i = idx - r * I;
    sIn[r * (I + 1) + i] = b0 + r < a.m ? __ldg(a.in + (b0 + r) * a.ld_in + static_cast<int64_t>(g) * a.in_group_stride + i) : 0.f;
  }
  __syncthreads();

  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  if (tx * 4 >= J || ty >= ty_n) return;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const float* in_rows = sIn + (ty * 4) * (I + 1);
  for (int i = 0; i < I; ++i) {
    const float4 m4 = *reinterpret_cast<const float4*>(sM + i * Jp + tx * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float v = in_rows[r * (I + 1) + i];
      acc[r][0] = fmaf(v, m4.x, acc[r][0]);
      acc[r][1] = fmaf(v, m4.y, acc[r][1]);
      acc[r][2] = fmaf(v, m4.z, acc[r][2]);
      acc[r][3] = fmaf(v, m4.w, acc[r][3]);
    }
  }
  const int j0 = tx * 4;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t b = b0 + ty * 4 + r;
    if (b >= a.m) continue;
    float* dst = a.out + b * a.ld_out + static_cast<int64_t>(g) * J + j0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (j0 + c < J) dst[c] = acc[r][c] + (a.bias ? __ldg(a.bias + g * J + j0 + c) : 0.f);
  }
}

// ------------------------------------------------------------------------------- weight gradient
// d_w[g][n][k] = sum_b dz[b, g, n] * in[b, g, k].  CTA = (row chunk, group); thread = one 4 x 4 block of
// the [N, K] gradient for one of `rs_n` interleaved row subsets; subsets are summed in order at the end
// and the per-chunk partials are added in chunk order by tower_wgrad_reduce_kernel (deterministic).
constexpr int kWgradTile = 64;

__global__ void __launch_bounds__(kThreads) tower_wgrad_kernel(const aread_tower_wgrad_args a, int mt_n, int rs_n,
                                                               int64_t rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const int N = a.n, K = a.k;
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  const int n4 = Np / 4, k4 = Kp / 4;
  float* sDz = smem;                       // [kWgradTile][Np]
  float* sIn = smem + kWgradTile * Np;     // [kWgradTile][Kp]
  const int g = blockIdx.y;
  const int mt = threadIdx.x % mt_n, rs = threadIdx.x / mt_n;
  const bool live = mt < n4 * k4 && rs < rs_n;
  const int nq = live ? mt / k4 : 0, kq = live ? mt - (mt / k4) * k4 : 0;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * rows_per_chunk;
  const int64_t c1 = min(a.m, c0 + rows_per_chunk);
  for (int64_t b0 = c0; b0 < c1; b0 += kWgradTile) {
    const int rows = c1 - b0 < kWgradTile ? static_cast<int>(c1 - b0) : kWgradTile;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kWgradTile * Np; idx += kThreads) {
      const int r = idx / Np, n = idx - r * Np;
      sDz[idx] = (r < rows && n < N) ? __ldg(a.dz + (b0 + r) * a.ld_dz + static_cast<int64_t>(g) * N + n) : 0.f;
    }
    for (int idx = threadIdx.x; idx < kWgradTile * Kp; idx += kThreads) {
      const int r = idx / Kp, k = idx - r * Kp;
      sIn[idx] = (r < rows && k < K) ? __ldg(a.in + (b0 + r) * a.ld_in + static_cast<int64_t>(g) * a.in_group_stride + k) : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int r = rs; r < rows; r += rs_n) {
        const float4 d = *reinterpret_cast<const float4*>(sDz + r * Np + nq * 4);
        const float4 x = *reinterpret_cast<const float4*>(sIn + r * Kp + kq * 4);
        acc[0][0] = fmaf(d.x, x.x, acc[0][0]); acc[0][1] = fmaf(d.x, x.y, acc[0][1]);
        acc[0][2] = fmaf(d.x, x.z, acc[0][2]); acc[0][3] = fmaf(d.x, x.w, acc[0][3]);
        acc[1][0] = fmaf(d.y, x.x, acc[1][0]); acc[1][1] = fmaf(d.y, x.y, acc[1][1]);
        acc[1][2] = fmaf(d.y, x.z, acc[1][2]); acc[1][3] = fmaf(d.y, x.w, acc[1][3]);
        acc[2][0] = fmaf(d.z, x.x, acc[2][0]); acc[2][1] = fmaf(d.z, x.y, acc[2][1]);
        acc[2][2] = fmaf(d.z, x.z, acc[2][2]); acc[2][3] = fmaf(d.z, x.w, acc[2][3]);
        acc[3][0] = fmaf(d.w, x.x, acc[3][0]); acc[3][1] = fmaf(d.w, x.y, acc[3][1]);
        acc[3][2] = fmaf(d.w, x.z, acc[3][2]); acc[3][3] = fmaf(d.w, x.w, acc[3][3]);
      }
    }
  }
  // combine the row subsets in order through shared memory (re-using the tile buffers)
  __syncthreads();
  float* sRed = smem;  // [rs_n][mt_n][16]
  if (live) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) sRed[(rs * mt_n + mt) * 16 + r * 4 + c] = acc[r][c];
  }
  __syncthreads();
  if (live && rs == 0) {
    float* dst = partial + (static_cast<int64_t>(blockIdx.x) * gridDim.y + g) * N * K;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = 0.f;
        for (int s = 0; s < rs_n; ++s) v += sRed[(s * mt_n + mt) * 16 + r * 4 + c];
        const int n = nq * 4 + r, k = kq * 4 + c;
        if (n < N && k < K) dst[n * K + k] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) tower_wgrad_reduce_kernel(int n_chunks, int64_t elems,
                                                                      const float* __restrict__ partial,
                                                                      float* __restrict__ d_w) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < elems;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < n_chunks; ++c) acc += partial[static_cast<int64_t>(c) * elems + i];
    d_w[i] = acc;
  }
}


// ------------------------------------------------------------------------------- gate mixing
// HEI gate of one level (model/aread.py:282-288): s = softmax(logits); under a HEMP mask
// sm = s * edge, r = sm / (sum sm + 1e-8), without a mask r = s; the tower's input is the r-weighted
// sum of the previous level's outputs.  Row-local.  A CTA stages kGateRows samples in shared memory with
// straight coalesced copies (every operand is a contiguous block of rows), one thread per (sample, tower)
// turns logits into weights in place, and the outputs are produced element by element, coalesced.
constexpr int kMaxPrev = 32;
constexpr int kGateRows = 32;

// in-place: lg[0..n_prev) holds logits on entry; on exit s[] = softmax (times nothing) and r[] the mixing weights
__device__ __forceinline__ void gate_weights(int n_prev, const float* __restrict__ edges_t, const float* lg, float* s,
                                             float* r, float& denom) {
  float mx = -INFINITY;
  for (int j = 0; j < n_prev; ++j) mx = fmaxf(mx, lg[j]);
  float sum = 0.f;
  for (int j = 0; j < n_prev; ++j) { const float e = expf(lg[j] - mx); s[j] = e; sum += e; }
  const float inv = 1.f / sum;
  denom = 1.f;
  if (edges_t != nullptr) {
    float tot = 0.f;
    for (int j = 0; j < n_prev; ++j) { const float v = s[j] * inv; s[j] = v; const float w = v * edges_t[j]; r[j] = w; tot += w; }
    denom = tot + 1e-8f;
    for (int j = 0; j < n_prev; ++j) r[j] = r[j] / denom;
  } else {
    for (int j = 0; j < n_prev; ++j) { const float v = s[j] * inv; s[j] = v; r[j] = v; }
  }
}

// Stage `n_valid` consecutive floats (whole rows of `w` elements) into shared memory with row stride wp >= w.
// Odd strides keep the per-(sample, tower) threads of the weight phase off each other's banks.
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int64_t n_valid, int n_total, int w,
                                           int wp) {
  for (int i = threadIdx.x; i < n_total; i += kThreads) {
    const int r = i / w, c = i - r * w;
    dst[r * wp + c] = i < n_valid ? __ldg(src + i) : 0.f;
  }
}

// logits of `rows` samples (row stride ld_logits, optional additive offset per logit) -> sS[(r * NT + t)][NPp]
__device__ __forceinline__ void stage_logits(float* sS, const aread_gate_mix_args& a, int64_t b0, int rows, int LP, int NP,
                                             int NPp) {
  const int64_t ld = a.ld_logits > 0 ? a.ld_logits : LP;
  for (int i = threadIdx.x; i < kGateRows * LP; i += kThreads) {
    const int r = i / LP, c = i - r * LP;
    float v = 0.f;
    if (r < rows) {
      v = __ldg(a.logits + (b0 + r) * ld + c);
      if (a.logit_offset != nullptr) v += __ldg(a.logit_offset + c);
    }
    const int q = i / NP;
    sS[q * NPp + (i - q * NP)] = v;
  }
}

__global__ void __launch_bounds__(kThreads) gate_mix_fwd_kernel(const aread_gate_mix_args a) {
  extern __shared__ float smem[];
  const int NT = a.n_tower, NP = a.n_prev, NA = a.n_prev_active, W = a.width;
  const int NPp = NP | 1, Wp = W | 1;
  const int LP = NT * NP, UP = NA * W, OP = NT * W;
  float* sS = smem;                            // [rows * NT][NPp]  logits -> softmax s
  float* sR = sS + kGateRows * NT * NPp;       // [rows * NT][NPp]  mixing weights r
  float* sU = sR + kGateRows * NT * NPp;       // [rows * NA][Wp]
  float* sE = sU + kGateRows * NA * Wp;        // [LP] edges (1 when unmasked)
  int* sSlot = reinterpret_cast<int*>(sE + LP);  // [NP]
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kGateRows;
  const int rows = a.m - b0 < kGateRows ? static_cast<int>(a.m - b0) : kGateRows;
  stage_logits(sS, a, b0, rows, LP, NP, NPp);
  if (UP > 0) stage_rows(sU, a.u_prev + b0 * UP, static_cast<int64_t>(rows) * UP, kGateRows * UP, W, Wp);
  for (int i = threadIdx.x; i < LP; i += kThreads) sE[i] = a.edges ? __ldg(a.edges + i) : 1.f;
  for (int i = threadIdx.x; i < NP; i += kThreads) sSlot[i] = __ldg(a.prev_slot + i);
  __syncthreads();
  for (int it = threadIdx.x; it < rows * NT; it += kThreads) {
    const int t = it % NT;
    float denom;
    gate_weights(NP, a.edges ? sE + t * NP : nullptr, sS + it * NPp, sS + it * NPp, sR + it * NPp, denom);
    if (a.sm != nullptr)
      for (int j = 0; j < NP; ++j) a.sm[(b0 * NT + it) * NP + j] = sS[it * NPp + j] * sE[t * NP + j];
  }
  __syncthreads();
  float* out = a.out + b0 * OP;
  for (int i = threadIdx.x; i < rows * OP; i += kThreads) {
    const int r = i / OP, rem = i - r * OP;
    const int t = rem / W, c = rem - t * W;
    const float* w = sR + (r * NT + t) * NPp;
    const float* u = sU + r * NA * Wp + c;
    float acc = 0.f;
    for (int j = 0; j < NP; ++j) {
      const int slot = sSlot[j];
      if (slot >= 0) acc = fmaf(w[j], u[slot * Wp], acc);
    }
    out[i] = acc;
  }
}

// d_logits[b, t, :] and d_u_prev[b, slot, :] = sum_t r[b, t, j(slot)] * d_out[b, t, :]   (towers in ascending order)
__global__ void __launch_bounds__(kThreads) gate_mix_bwd_kernel(const aread_gate_mix_args a) {
  extern __shared__ float smem[];
  const int NT = a.n_tower, NP = a.n_prev, NA = a.n_prev_active, W = a.width;
  const int NPp = NP | 1, Wp = W | 1;
  const int LP = NT * NP, UP = NA * W, OP = NT * W;
  float* sS = smem;                            // [rows * NT][NPp]  logits -> s -> d_logits
  float* sR = sS + kGateRows * NT * NPp;       // r
  float* sD = sR + kGateRows * NT * NPp;       // dr
  float* sU = sD + kGateRows * NT * NPp;       // [rows * NA][Wp]
  float* sG = sU + kGateRows * NA * Wp;        // [rows * NT][Wp]  d_out
  float* sE = sG + kGateRows * NT * Wp;        // [LP]
  int* sSlot = reinterpret_cast<int*>(sE + LP);  // [NP] then [NA]
  int* sTower = sSlot + NP;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kGateRows;
  const int rows = a.m - b0 < kGateRows ? static_cast<int>(a.m - b0) : kGateRows;
  stage_logits(sS, a, b0, rows, LP, NP, NPp);
  if (UP > 0) stage_rows(sU, a.u_prev + b0 * UP, static_cast<int64_t>(rows) * UP, kGateRows * UP, W, Wp);
  stage_rows(sG, a.d_out + b0 * OP, static_cast<int64_t>(rows) * OP, kGateRows * OP, W, Wp);
  for (int i = threadIdx.x; i < LP; i += kThreads) sE[i] = a.edges ? __ldg(a.edges + i) : 1.f;
  for (int i = threadIdx.x; i < NP; i += kThreads) sSlot[i] = __ldg(a.prev_slot + i);
  for (int i = threadIdx.x; i < NA; i += kThreads) sTower[i] = __ldg(a.slot_tower + i);
  __syncthreads();
  for (int it = threadIdx.x; it < rows * NT; it += kThreads) {
    const int r = it / NT, t = it - r * NT;
    float* s = sS + it * NPp;
    float* w = sR + it * NPp;
    float* dr = sD + it * NPp;
    float denom;
    gate_weights(NP, a.edges ? sE + t * NP : nullptr, s, s, w, denom);
    const float* g = sG + it * Wp;
    float dot_r = 0.f;
    for (int j = 0; j < NP; ++j) {
      const int slot = sSlot[j];
      float acc = 0.f;
      if (slot >= 0) {
        const float* u = sU + (r * NA + slot) * Wp;
        for (int c = 0; c < W; ++c) acc = fmaf(g[c], u[c], acc);
      }
      dr[j] = acc;
      dot_r = fmaf(acc, w[j], dot_r);
    }
    float dot_s = 0.f;
    for (int j = 0; j < NP; ++j) {  // dr -> ds (through the renormalisation and the mask)
      if (a.edges != nullptr) dr[j] = (dr[j] - dot_r) / denom * sE[t * NP + j];
      dot_s = fmaf(s[j], dr[j], dot_s);
    }
    for (int j = 0; j < NP; ++j) s[j] = s[j] * (dr[j] - dot_s);
  }
  __syncthreads();
  {
    const int64_t ldd = a.ld_dlogits > 0 ? a.ld_dlogits : LP;
    for (int i = threadIdx.x; i < rows * LP; i += kThreads) {
      const int q = i / NP, r = i / LP;
      a.d_logits[(b0 + r) * ldd + (i - r * LP)] = sS[q * NPp + (i - q * NP)];
    }
  }
  if (UP > 0 && a.d_u_prev != nullptr) {
    float* d_u = a.d_u_prev + b0 * UP;
    for (int i = threadIdx.x; i < rows * UP; i += kThreads) {
      const int r = i / UP, rem = i - r * UP;
      const int slot = rem / W, c = rem - slot * W;
      const int j = sTower[slot];
      const float* w = sR + r * NT * NPp + j;
      const float* g = sG + r * NT * Wp + c;
      float acc = 0.f;
      for (int t = 0; t < NT; ++t) acc = fmaf(w[t * NPp], g[t * Wp], acc);
      d_u[i] = acc;
    }
  }
}

// ---- direct variants for the shipped sizes (n_prev <= 8, width % 4 == 0): no staging.  Forward: one thread per
// (sample, tower) keeps the gate weights in registers and walks its output row in float4s; the previous level's
// activations of a sample are read by all its towers' threads (neighbours in the warp: L1 broadcasts).  Backward:
// a CTA owns 256 / n_tower samples; phase 1 (thread = sample, tower) turns d_out into d_logits and leaves the mixing
// weights in shared memory, phase 2 (thread = sample, previous tower, float4) gathers d_u_prev in tower order.
constexpr int kDirectPrev = 8;

struct GateRow {
  float s[kDirectPrev], r[kDirectPrev], e[kDirectPrev];
  float denom;
};

__device__ __forceinline__ void gate_row(const aread_gate_mix_args& a, int64_t b, int t, GateRow& w) {
  const int NP = a.n_prev;
  const int64_t ld = a.ld_logits > 0 ? a.ld_logits : static_cast<int64_t>(a.n_tower) * NP;
  const float* lg = a.logits + b * ld + t * NP;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kDirectPrev; ++j)
    if (j < NP) {
      w.s[j] = __ldg(lg + j) + (a.logit_offset ? __ldg(a.logit_offset + t * NP + j) : 0.f);
      w.e[j] = a.edges ? __ldg(a.edges + t * NP + j) : 1.f;
      mx = fmaxf(mx, w.s[j]);
    }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kDirectPrev; ++j)
    if (j < NP) { w.s[j] = expf(w.s[j] - mx); sum += w.s[j]; }
  const float inv = 1.f / sum;
  w.denom = 1.f;
  if (a.edges != nullptr) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < kDirectPrev; ++j)
      if (j < NP) { w.s[j] *= inv; w.r[j] = w.s[j] * w.e[j]; tot += w.r[j]; }
    w.denom = tot + 1e-8f;
#pragma unroll
    for (int j = 0; j < kDirectPrev; ++j)
      if (j < NP) w.r[j] = w.r[j] / w.denom;
  } else {
#pragma unroll
    for (int j = 0; j < kDirectPrev; ++j)
      if (j < NP) { w.s[j] *= inv; w.r[j] = w.s[j]; }
  }
}

// Forward in two phases per tile of rows, so that a row of u is read once and not once per tower:
//   1. thread = (row, tower): softmax / edge renormalisation of its logits -> r[row][tower][prev] in shared memory
//   2. thread = (row, 4 columns, share of the towers): the previous level's active outputs of those columns in
//      registers, then out[row][tower][columns] = sum_prev r * u for its towers (128-bit loads and stores)
__global__ void __launch_bounds__(kThreads) gate_mix_fwd_tile_kernel(const aread_gate_mix_args a, int rows_per_cta, int ts) {
  extern __shared__ float s_r[];                     // [rows_per_cta * NT][NP]
  __shared__ int s_slot[kDirectPrev];
  const int NT = a.n_tower, NP = a.n_prev, NA = a.n_prev_active, W = a.width, W4 = a.width / 4;
  if (threadIdx.x < kDirectPrev) s_slot[threadIdx.x] = threadIdx.x < NP ? a.prev_slot[threadIdx.x] : -1;
  const int64_t n_tiles = (a.m + rows_per_cta - 1) / rows_per_cta;
  const int t_per = (NT + ts - 1) / ts;              // towers per phase-2 thread
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * rows_per_cta;
    const int rows = a.m - b0 < rows_per_cta ? static_cast<int>(a.m - b0) : rows_per_cta;
    __syncthreads();
    if (threadIdx.x < rows * NT) {
      const int r = threadIdx.x / NT, t = threadIdx.x - r * NT;
      const int64_t i = (b0 + r) * NT + t;
      GateRow w;
      gate_row(a, b0 + r, t, w);
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j)
        if (j < NP) {
          s_r[threadIdx.x * NP + j] = w.r[j];
          if (a.sm != nullptr) a.sm[i * NP + j] = w.s[j] * w.e[j];
        }
    }
    __syncthreads();
    const int items = rows * W4 * ts;
    for (int it = threadIdx.x; it < items; it += kThreads) {
      const int r = it / (W4 * ts), rem = it - r * (W4 * ts);
      const int part = rem / W4, c = rem - part * W4;
      const int64_t b = b0 + r;
      const float4* u = reinterpret_cast<const float4*>(a.u_prev + b * (static_cast<int64_t>(NA) * W)) + c;
      float4 uu[kDirectPrev];
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j) {
        const int slot = s_slot[j];
        uu[j] = slot >= 0 ? __ldg(u + slot * W4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int t1 = (part + 1) * t_per < NT ? (part + 1) * t_per : NT;
      float4* out = reinterpret_cast<float4*>(a.out + b * (static_cast<int64_t>(NT) * W)) + c;
      for (int t = part * t_per; t < t1; ++t) {
        const float* rr = s_r + (r * NT + t) * NP;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kDirectPrev; ++j)
          if (j < NP) {
            const float wj = rr[j];
            acc.x = fmaf(wj, uu[j].x, acc.x); acc.y = fmaf(wj, uu[j].y, acc.y);
            acc.z = fmaf(wj, uu[j].z, acc.z); acc.w = fmaf(wj, uu[j].w, acc.w);
          }
        out[t * W4] = acc;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) gate_mix_bwd_direct_kernel(const aread_gate_mix_args a, int rows_per_cta) {
  extern __shared__ float s_r[];                     // [rows_per_cta * NT][NP], then u of the tile [rows][NA * W + 4]
  const int NT = a.n_tower, NP = a.n_prev, NA = a.n_prev_active, W = a.width, W4 = a.width / 4;
  const int ldu = NA * W + 4;                        // padded row: the rows of a warp fall on different banks
  float* s_u = s_r + ((rows_per_cta * NT * NP + 3) & ~3);
  const int64_t ldd = a.ld_dlogits > 0 ? a.ld_dlogits : static_cast<int64_t>(NT) * NP;
  const int64_t n_tiles = (a.m + rows_per_cta - 1) / rows_per_cta;
  int slots[kDirectPrev];
#pragma unroll
  for (int j = 0; j < kDirectPrev; ++j) slots[j] = j < NP ? a.prev_slot[j] : -1;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * rows_per_cta;
    const int rows = a.m - b0 < rows_per_cta ? static_cast<int>(a.m - b0) : rows_per_cta;
    __syncthreads();
    // u of the tile, read once (the NT threads of a row all need it), whole rows at a time
    for (int it = threadIdx.x; it < rows * NA * W4; it += kThreads) {
      const int r = it / (NA * W4), q = it - r * (NA * W4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.u_prev + (b0 + r) * (static_cast<int64_t>(NA) * W)) + q);
      *reinterpret_cast<float4*>(s_u + r * ldu + q * 4) = v;
    }
    __syncthreads();
    if (threadIdx.x < rows * NT) {
      const int r = threadIdx.x / NT, t = threadIdx.x - r * NT;
      const int64_t b = b0 + r;
      GateRow w;
      gate_row(a, b, t, w);
      const float* u = s_u + r * ldu;
      const float4* g = reinterpret_cast<const float4*>(a.d_out + (b * NT + t) * W);
      float dr[kDirectPrev];
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j) dr[j] = 0.f;
      for (int c = 0; c < W4; ++c) {
        const float4 gv = __ldg(g + c);
#pragma unroll
        for (int j = 0; j < kDirectPrev; ++j) {
          const int slot = slots[j];
          if (slot >= 0) {
            const float4 v = *reinterpret_cast<const float4*>(u + slot * W + c * 4);
            dr[j] = fmaf(gv.x, v.x, fmaf(gv.y, v.y, fmaf(gv.z, v.z, fmaf(gv.w, v.w, dr[j]))));
          }
        }
      }
      float dot_r = 0.f;
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j)
        if (j < NP) dot_r = fmaf(dr[j], w.r[j], dot_r);
      float dot_s = 0.f;
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j)
        if (j < NP) {
          if (a.edges != nullptr) dr[j] = (dr[j] - dot_r) / w.denom * w.e[j];
          dot_s = fmaf(w.s[j], dr[j], dot_s);
        }
      float* dl = a.d_logits + b * ldd + t * NP;
#pragma unroll
      for (int j = 0; j < kDirectPrev; ++j)
        if (j < NP) {
          dl[j] = w.s[j] * (dr[j] - dot_s);
          s_r[threadIdx.x * NP + j] = w.r[j];
        }
    }
    __syncthreads();
    if (a.d_u_prev != nullptr) {
      const int items = rows * NA * W4;
      for (int it = threadIdx.x; it < items; it += kThreads) {
        const int r = it / (NA * W4), rem = it - r * (NA * W4);
        const int slot = rem / W4, c = rem - slot * W4;
        const int j = __ldg(a.slot_tower + slot);
        const int64_t b = b0 + r;
        const float4* g = reinterpret_cast<const float4*>(a.d_out + b * (static_cast<int64_t>(NT) * W)) + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < NT; ++t) {
          const float wt = s_r[(r * NT + t) * NP + j];
          const float4 gv = __ldg(g + t * W4);
          acc.x = fmaf(wt, gv.x, acc.x); acc.y = fmaf(wt, gv.y, acc.y);
          acc.z = fmaf(wt, gv.z, acc.z); acc.w = fmaf(wt, gv.w, acc.w);
        }
        reinterpret_cast<float4*>(a.d_u_prev + b * (static_cast<int64_t>(NA) * W))[slot * W4 + c] = acc;
      }
    }
  }
}

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int wgrad_chunks(int64_t m, int groups) {
  int64_t c = (m + 4 * kWgradTile - 1) / (4 * kWgradTile);  // at least 256 rows per chunk
  const int64_t cap = (2 * kNumSMs + groups - 1) / groups;
  if (c > cap) c = cap;
  return static_cast<int>(c < 1 ? 1 : c);
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_tower_linear(const aread_tower_linear_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "tower_linear: null args");
  const aread_tower_linear_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.groups > 0 && a.in_width > 0 && a.out_width > 0, "tower_linear: bad shape");
  AREAD_REQUIRE(a.out_width <= 128 && a.in_width <= 256, "tower_linear: layer %d -> %d is too wide for this kernel",
                a.in_width, a.out_width);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.in && a.out && a.weight, "tower_linear: null pointer");
  const int j4 = (a.out_width + 3) / 4;
  const int tx_n = pow2_ceil(j4);
  int ty_n = kThreads / tx_n;
  if (ty_n > 32) ty_n = 32;  // at most 128 rows per tile
  const int tile_rows = ty_n * 4;
  const int jp = (a.out_width + 3) & ~3;
  const size_t smem = sizeof(float) * (static_cast<size_t>(a.in_width) * jp + static_cast<size_t>(tile_rows) * (a.in_width + 1));
  AREAD_REQUIRE(smem <= 200 * 1024, "tower_linear: tile does not fit shared memory");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (smem > 48 * 1024)
    AREAD_CUDA(cudaFuncSetAttribute(tower_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const dim3 grid(static_cast<unsigned>((a.m + tile_rows - 1) / tile_rows), static_cast<unsigned>(a.groups));
  AREAD_LAUNCH(tower_linear_kernel, grid, kThreads, smem, stream, a, tx_n, ty_n);
  return AREAD_OK;
}

size_t aread_tower_wgrad_workspace_bytes(int64_t m, int32_t groups, int32_t n, int32_t k) {
  return aread::align_up(static_cast<size_t>(aread::wgrad_chunks(m, groups)) * groups * n * k * 4, 256);
}

int aread_tower_wgrad(const aread_tower_wgrad_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "tower_wgrad: null args");
  const aread_tower_wgrad_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.groups > 0 && a.n > 0 && a.k > 0, "tower_wgrad: bad shape");
  AREAD_REQUIRE(a.n <= 128 && a.k <= 256, "tower_wgrad: layer %d -> %d is too wide for this kernel", a.k, a.n);
  AREAD_REQUIRE(a.d_w && a.workspace, "tower_wgrad: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t elems = static_cast<int64_t>(a.groups) * a.n * a.k;
  if (a.m == 0) {
    AREAD_CUDA(cudaMemsetAsync(a.d_w, 0, elems * 4, stream));
    return AREAD_OK;
  }
  AREAD_REQUIRE(a.dz && a.in, "tower_wgrad: null pointer");
  const int chunks = wgrad_chunks(a.m, a.groups);
  AREAD_REQUIRE(a.workspace_bytes >= static_cast<size_t>(chunks) * elems * 4, "tower_wgrad: workspace too small");
  const int np = (a.n + 3) & ~3, kp = (a.k + 3) & ~3;
  const int blocks = (np / 4) * (kp / 4);
  AREAD_REQUIRE(blocks <= kThreads, "tower_wgrad: %d x %d gradient is too large for this kernel (n * k <= 4096)", a.n,
                a.k);
  const int mt_n = pow2_ceil(blocks);
  const int rs_n = kThreads / mt_n;
  size_t smem = sizeof(float) * kWgradTile * (np + kp);
  const size_t red = sizeof(float) * static_cast<size_t>(rs_n) * mt_n * 16;
  if (red > smem) smem = red;
  if (smem > 48 * 1024)
    AREAD_CUDA(cudaFuncSetAttribute(tower_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t rows_per_chunk = (a.m + chunks - 1) / chunks;
  float* partial = static_cast<float*>(a.workspace);
  const dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(a.groups));
  AREAD_LAUNCH(tower_wgrad_kernel, grid, kThreads, smem, stream, a, mt_n, rs_n, rows_per_chunk, partial);
  AREAD_LAUNCH(tower_wgrad_reduce_kernel, static_cast<unsigned>((elems + kThreads - 1) / kThreads), kThreads, 0, stream,
               chunks, elems, partial, a.d_w);
  return AREAD_OK;
}

}  // extern "C"

extern "C" int aread_gate_mix(const aread_gate_mix_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "gate_mix: null args");
  const aread_gate_mix_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.n_tower > 0 && a.n_prev > 0 && a.n_prev <= kMaxPrev && a.n_prev_active >= 0 && a.width > 0,
                "gate_mix: bad shape (n_prev %d, max %d)", a.n_prev, kMaxPrev);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.logits && a.prev_slot && (a.u_prev || a.n_prev_active == 0), "gate_mix: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t lp = static_cast<size_t>(a.n_tower) * a.n_prev;
  const size_t ltp = static_cast<size_t>(a.n_tower) * (a.n_prev | 1), utp = static_cast<size_t>(a.n_prev_active) * (a.width | 1),
               otp = static_cast<size_t>(a.n_tower) * (a.width | 1);   // padded shared-memory rows
  const unsigned grid = static_cast<unsigned>((a.m + kGateRows - 1) / kGateRows);
  static uint64_t configured = 0;          // cudaFuncSetAttribute is per device
  if (first_use_on_device(&configured)) {
    AREAD_CUDA(cudaFuncSetAttribute(gate_mix_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AREAD_CUDA(cudaFuncSetAttribute(gate_mix_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const bool direct = a.n_prev <= kDirectPrev && a.width % 4 == 0 && a.n_tower <= kThreads && a.n_prev_active > 0 &&
                      ((reinterpret_cast<uintptr_t>(a.u_prev) | reinterpret_cast<uintptr_t>(a.out) |
                        reinterpret_cast<uintptr_t>(a.d_out) | reinterpret_cast<uintptr_t>(a.d_u_prev)) & 15) == 0;
  if (direct && a.d_out == nullptr) {
    AREAD_REQUIRE(a.out != nullptr, "gate_mix: null out");
    const int rows_per_cta = kThreads / a.n_tower;
    const int w4 = a.width / 4;
    int ts = kThreads / (rows_per_cta * w4);           // tower shares so that phase 2 has a thread's worth of work each
    if (ts < 1) ts = 1;
    if (ts > a.n_tower) ts = a.n_tower;
    int64_t g = (a.m + rows_per_cta - 1) / rows_per_cta;
    if (g > kNumSMs * 16) g = kNumSMs * 16;
    AREAD_LAUNCH(gate_mix_fwd_tile_kernel, static_cast<unsigned>(g), kThreads,
                 sizeof(float) * rows_per_cta * a.n_tower * a.n_prev, stream, a, rows_per_cta, ts);
    return AREAD_OK;
  }
  if (direct) {
    AREAD_REQUIRE(a.d_logits && a.slot_tower, "gate_mix: null gradient pointer");
    int rows_per_cta = kThreads / a.n_tower;
    const size_t per_row = sizeof(float) * (static_cast<size_t>(a.n_tower) * a.n_prev + a.n_prev_active * a.width + 4);
    if (rows_per_cta * per_row + 16 > 40 * 1024) rows_per_cta = static_cast<int>((40 * 1024 - 16) / per_row);
    AREAD_REQUIRE(rows_per_cta >= 1, "gate_mix: one row needs %zu bytes of shared memory", per_row);
    int64_t g = (a.m + rows_per_cta - 1) / rows_per_cta;
    if (g > kNumSMs * 16) g = kNumSMs * 16;
    const size_t smem = sizeof(float) * (((static_cast<size_t>(rows_per_cta) * a.n_tower * a.n_prev + 3) & ~size_t(3)) +
                                         static_cast<size_t>(rows_per_cta) * (a.n_prev_active * a.width + 4));
    AREAD_LAUNCH(gate_mix_bwd_direct_kernel, static_cast<unsigned>(g), kThreads, smem, stream, a, rows_per_cta);
    return AREAD_OK;
  }
  if (a.d_out == nullptr) {
    AREAD_REQUIRE(a.out != nullptr, "gate_mix: null out");
    const size_t smem = sizeof(float) * (kGateRows * (2 * ltp + utp) + lp + a.n_prev);
    AREAD_REQUIRE(smem <= 200 * 1024, "gate_mix: level too wide for shared memory (%zu bytes)", smem);
    AREAD_LAUNCH(gate_mix_fwd_kernel, grid, kThreads, smem, stream, a);
  } else {
    AREAD_REQUIRE(a.d_logits && (a.d_u_prev || a.n_prev_active == 0) && a.slot_tower, "gate_mix: null gradient pointer");
    const size_t smem = sizeof(float) * (kGateRows * (3 * ltp + utp + otp) + lp + a.n_prev + a.n_prev_active);
    AREAD_REQUIRE(smem <= 200 * 1024, "gate_mix: level too wide for shared memory (%zu bytes)", smem);
    AREAD_LAUNCH(gate_mix_bwd_kernel, grid, kThreads, smem, stream, a);
  }
  return AREAD_OK;
}
