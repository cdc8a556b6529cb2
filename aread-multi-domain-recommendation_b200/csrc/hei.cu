// One HEI tower layer (Linear -> BatchNorm1d -> ReLU -> Dropout, model/layer.py:221-229) for all towers of a
// level that run under the current HEMP mask (model/aread.py:297-321), with the passes around the small
// Linear fused into it.  Widths are 8..64, far below a tensor-core tile: this is fp32 CUDA-core work
// bound by how often the [m, groups * width] activations cross HBM / L2.
//
// forward  (aread_hei_layer_fwd):  reads the layer input ONCE -- either a plain activation or the previous
//   layer's pre-activation z, normalised / rectified / dropped on the fly -- multiplies by W, writes z and
//   accumulates the BatchNorm column sums of z in the same pass; a tiny second kernel finalises the
//   statistics.  The activation tensor between two layers is never materialised.
// backward (aread_hei_layer_bwd):  one pass computes dz from (z, d_out, statistics), the weight gradient
//   dz^T x (x recomputed from the previous pre-activation), the input gradient dz W, and -- when the input is
//   itself a BatchNorm'd layer -- the column sums its BatchNorm backward needs.
// Replaces per layer: tower_linear + bn_stats + bn_act (forward) and bn_bwd_stats + bn_bwd_apply +
// tower_wgrad + tower_linear (backward).  All reductions run in a fixed order (bit-reproducible).
#include <cstdlib>

#include "bn_common.cuh"
#include "common.cuh"
#include "tensor_map.cuh"
#include "hei_tc.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxWidth = 64;     // k, n <= 64 (tower_dims of every shipped config); wider layers use tower.cu
constexpr int kMaxTileQuads = 64; // row tile <= 256 rows

struct Drop {
  uint32_t threshold;
  float keep_scale;
};
inline Drop make_drop(float p) {
  Drop d;
  d.threshold = p > 0.f ? dropout_threshold(p) : 0u;
  d.keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return d;
}

// element of the layer input: plain, or dropout(relu(bn(z_prev))) recomputed from the pre-activation
__device__ __forceinline__ float src_value(float raw, const float* __restrict__ scale, const float* __restrict__ shift,
                                           int col, uint64_t seed, uint32_t salt, uint64_t flat, uint32_t thr,
                                           float keep_scale) {
  if (scale == nullptr) return raw;
  const bool keep = thr == 0u || dropout_keep(seed, salt, flat, thr);
  return act_value(raw, __ldg(scale + col), __ldg(shift + col), keep, keep_scale);
}

// ------------------------------------------------------------------------------------------ forward
// CTA = (chunk of row tiles, group).  Thread = 4 rows x 4 output columns of the tile; the reduction dimension is
// walked four at a time with 128-bit shared-memory reads (rows are padded by 4 floats: conflict-free per
// quarter warp).
__global__ void __launch_bounds__(kThreads, 3) hei_layer_fwd_kernel(const aread_hei_layer_fwd_args a, int tx_n, int ty_n,
                                                                 int tiles_per_cta, uint32_t thr, float keep_scale,
                                                                 int do_stats, float* __restrict__ partial) {
  const uint64_t seed = seed_of(a);
  extern __shared__ __align__(16) float smem[];
  const int K = a.k, N = a.n, G = a.groups;
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  const int k4 = Kp / 4;
  const int tile_rows = ty_n * 4;
  const int ldi = Kp + 4;
  float* sM = smem;              // [Kp][Np]  W^T, zero padded
  float* sPivot = sM + Kp * Np;  // [Np]
  float* sIn = sPivot + Np;      // [tile_rows][Kp + 4]; afterwards reduction scratch [ty_n][2][Np]
  const int g = blockIdx.y;
  const int src_width = G * K, out_width = G * N;
  const float* __restrict__ w = a.weight + static_cast<int64_t>(g) * N * K;
  const float* __restrict__ bias = a.bias ? a.bias + g * N : nullptr;
  const float* __restrict__ sscale = a.src_scale;
  const float* __restrict__ sshift = a.src_shift;

  for (int idx = threadIdx.x; idx < Kp * Np; idx += kThreads) {
    const int i = idx / Np, j = idx - i * Np;
    sM[idx] = (i < K && j < N) ? __ldg(w + static_cast<int64_t>(j) * K + i) : 0.f;
  }
  if (do_stats) {  // z of row 0, same arithmetic as below: the pivot of the variance sums (bn_finalize reads z[0])
    for (int i = threadIdx.x; i < Kp; i += kThreads) {
      const int col = g * K + i;
      sIn[i] = i < K ? src_value(__ldg(a.src + col), sscale, sshift, col, seed, a.src_salt, static_cast<uint64_t>(col),
                                 thr, keep_scale)
                     : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < Np) {
      float acc = 0.f;
      for (int i = 0; i < Kp; ++i) acc = fmaf(sIn[i], sM[i * Np + threadIdx.x], acc);
      sPivot[threadIdx.x] = acc + ((bias && threadIdx.x < N) ? __ldg(bias + threadIdx.x) : 0.f);
    }
  }

  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const bool live = ty < ty_n;
  const int j0 = tx * 4;
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  float bj[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) bj[c] = (bias && j0 + c < N) ? __ldg(bias + j0 + c) : 0.f;
  const bool vec_out = (N % 4 == 0);
  const bool vec_in = (K % 4 == 0) && (a.ld_src % 4 == 0) && (reinterpret_cast<uintptr_t>(a.src) % 16 == 0);
  // loader role: one quad of input columns, rows lq, lq + l_rows, ...
  const int lk = threadIdx.x % k4, lq = threadIdx.x / k4, l_rows = kThreads / k4;
  const int lcol = g * K + lk * 4;
  float lsc[4] = {0.f, 0.f, 0.f, 0.f}, lsh[4] = {0.f, 0.f, 0.f, 0.f};
  if (sscale) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (lk * 4 + c < K) { lsc[c] = __ldg(sscale + lcol + c); lsh[c] = __ldg(sshift + lcol + c); }
  }

  const int64_t n_tiles = (a.m + tile_rows - 1) / tile_rows;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < n_tiles ? t0 + tiles_per_cta : n_tiles;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t b0 = t * tile_rows;
    __syncthreads();
    if (lq < l_rows) {
      for (int r = lq; r < tile_rows; r += l_rows) {
        const int64_t row = b0 + r;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row < a.m) {
          const float* srow = a.src + row * a.ld_src + lcol;
          if (vec_in) {
            const float4 raw = ldg4(srow);
            v[0] = raw.x; v[1] = raw.y; v[2] = raw.z; v[3] = raw.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (lk * 4 + c < K) v[c] = __ldg(srow + c);
          }
          if (sscale) {
            const uint64_t flat = static_cast<uint64_t>(row) * src_width + lcol;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const bool keep = thr == 0u || dropout_keep(seed, a.src_salt, flat + c, thr);
              v[c] = (lk * 4 + c < K) ? act_value(v[c], lsc[c], lsh[c], keep, keep_scale) : 0.f;
            }
          }
        }
        *reinterpret_cast<float4*>(sIn + r * ldi + lk * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
    if (!live) continue;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* in_rows = sIn + (ty * 4) * ldi;
    for (int i0 = 0; i0 < Kp; i0 += 4) {
      float4 x4[4], m4[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) x4[r] = *reinterpret_cast<const float4*>(in_rows + r * ldi + i0);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) m4[ii] = *reinterpret_cast<const float4*>(sM + (i0 + ii) * Np + j0);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float xv[4] = {x4[r].x, x4[r].y, x4[r].z, x4[r].w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          acc[r][0] = fmaf(xv[ii], m4[ii].x, acc[r][0]);
          acc[r][1] = fmaf(xv[ii], m4[ii].y, acc[r][1]);
          acc[r][2] = fmaf(xv[ii], m4[ii].z, acc[r][2]);
          acc[r][3] = fmaf(xv[ii], m4[ii].w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t row = b0 + ty * 4 + r;
      if (row >= a.m) continue;
      float zv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) zv[c] = acc[r][c] + bj[c];
      float* dst = a.z + row * out_width + g * N + j0;
      if (vec_out) {
        *reinterpret_cast<float4*>(dst) = make_float4(zv[0], zv[1], zv[2], zv[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (j0 + c < N) dst[c] = zv[c];
      }
      if (do_stats) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float dv = zv[c] - sPivot[j0 + c];
          cs[c] += dv;
          cq[c] = fmaf(dv, dv, cq[c]);
        }
      }
    }
  }
  if (!do_stats) return;
  __syncthreads();
  float* sRed = sIn;  // [ty_n][2][Np]
  if (live) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sRed[(ty * 2 + 0) * Np + j0 + c] = cs[c];
      sRed[(ty * 2 + 1) * Np + j0 + c] = cq[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = 0.f, q = 0.f;
    for (int y = 0; y < ty_n; ++y) {
      s += sRed[(y * 2 + 0) * Np + threadIdx.x];
      q += sRed[(y * 2 + 1) * Np + threadIdx.x];
    }
    partial[(static_cast<int64_t>(blockIdx.x) * 2 + 0) * out_width + g * N + threadIdx.x] = s;
    partial[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * out_width + g * N + threadIdx.x] = q;
  }
}

// ------------------------------------------------------------------------------------------ backward
// CTA = (chunk of row tiles, group).  Per tile: dz -> smem, x (and z_prev) -> smem (both by threads that own a
// fixed quad of columns, 128-bit global loads), then every thread adds its share of dz^T x (a 4 x 4 block of
// [N, K] over an interleaved row subset) and computes 4 rows x 4 columns of d_in = dz W together with the
// BatchNorm-backward sums of the layer below.
__global__ void __launch_bounds__(kThreads, 3) hei_layer_bwd_kernel(const aread_hei_layer_bwd_args a, int tx_n, int ty_n,
                                                                    int mt_n, int rs_n, int tiles_per_cta, uint32_t thr,
                                                                    float keep_scale, uint32_t src_thr,
                                                                    float src_keep_scale, float* __restrict__ partial_w,
                                                                    float* __restrict__ partial_s) {
  const uint64_t seed = seed_of(a);
  extern __shared__ __align__(16) float smem[];
  const int K = a.k, N = a.n, G = a.groups;
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  const int n4 = Np / 4, k4 = Kp / 4;
  const int tile_rows = ty_n * 4;
  const int ldz = Np + 4, ldx = Kp + 4;
  const bool src_bn = a.src_scale != nullptr;
  float* sW = smem;                      // [Np][Kp], zero padded
  float* sDz = sW + Np * Kp;             // [tile_rows][Np + 4]
  float* sX = sDz + tile_rows * ldz;     // [tile_rows][Kp + 4]
  float* sZp = sX + tile_rows * ldx;     // [tile_rows][Kp + 4] (only when src_bn)
  const int g = blockIdx.y;
  const int width = G * N, src_width = G * K;
  const float* __restrict__ w = a.weight + static_cast<int64_t>(g) * N * K;

  for (int idx = threadIdx.x; idx < Np * Kp; idx += kThreads) {
    const int n = idx / Kp, k = idx - n * Kp;
    sW[idx] = (n < N && k < K) ? __ldg(w + static_cast<int64_t>(n) * K + k) : 0.f;
  }

  // weight-gradient role
  const int mt = threadIdx.x % mt_n, rs = threadIdx.x / mt_n;
  const bool w_live = rs < rs_n;
  const int nq = mt / k4, kq = mt - nq * k4;
  float aw[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) aw[i][j] = 0.f;
  // input-gradient role
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const bool d_live = ty < ty_n;
  const int c0 = tx * 4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  // loader roles: a fixed quad of columns each, rows strided
  const int an = threadIdx.x % n4, ar = threadIdx.x / n4, a_rows = kThreads / n4;
  const int acol = g * N + an * 4;
  const int bk = threadIdx.x % k4, br = threadIdx.x / k4, b_rows = kThreads / k4;
  const int bcol = g * K + bk * 4;
  const bool vec_n = (N % 4 == 0);
  const bool vec_k = (K % 4 == 0) && (a.ld_src % 4 == 0) && (reinterpret_cast<uintptr_t>(a.src) % 16 == 0);
  const bool vec_din = (K % 4 == 0);

  const int64_t n_tiles = (a.m + tile_rows - 1) / tile_rows;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * tiles_per_cta;
  const int64_t t1 = t0 + tiles_per_cta < n_tiles ? t0 + tiles_per_cta : n_tiles;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t b0 = t * tile_rows;
    const int rows = a.m - b0 < tile_rows ? static_cast<int>(a.m - b0) : tile_rows;
    __syncthreads();
    // dz = A * dy - B - (z - mean) * C  with A = scale, B = scale * mean(dy), C = scale * rstd * mean(dy * xhat);
    // dy = d_out * [y > 0] * keep / (1 - p)
    if (ar < a_rows) {
      float pa[4], psh[4], pb[4], pmu[4], pc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool ok = an * 4 + c < N;
        const float sc = ok ? __ldg(a.scale + acol + c) : 0.f;
        pa[c] = sc;
        psh[c] = ok ? __ldg(a.shift + acol + c) : 0.f;
        pmu[c] = ok ? __ldg(a.mean + acol + c) : 0.f;
        pb[c] = ok ? sc * __ldg(a.coef + acol + c) : 0.f;
        pc[c] = ok ? sc * __ldg(a.rstd + acol + c) * __ldg(a.coef + width + acol + c) : 0.f;
      }
      for (int r = ar; r < tile_rows; r += a_rows) {
        float dz[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
          const int64_t row = b0 + r;
          float z[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
          if (vec_n) {
            const float4 zz = ldg4(a.z + row * width + acol), dd = ldg4(a.d_out + row * width + acol);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
            d[0] = dd.x; d[1] = dd.y; d[2] = dd.z; d[3] = dd.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (an * 4 + c < N) { z[c] = __ldg(a.z + row * width + acol + c); d[c] = __ldg(a.d_out + row * width + acol + c); }
          }
          const uint64_t flat = static_cast<uint64_t>(row) * width + acol;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float y = fmaf(z[c], pa[c], psh[c]);
            const bool keep = thr == 0u || dropout_keep(seed, a.salt, flat + c, thr);
            const float dy = (y > 0.f && keep) ? d[c] * keep_scale : 0.f;
            dz[c] = a.bn_skip ? dy : fmaf(pa[c], dy, -pb[c]) - (z[c] - pmu[c]) * pc[c];
            if (an * 4 + c >= N) dz[c] = 0.f;
          }
        }
        *reinterpret_cast<float4*>(sDz + r * ldz + an * 4) = make_float4(dz[0], dz[1], dz[2], dz[3]);
      }
    }
    if (br < b_rows) {
      float qs[4] = {0.f, 0.f, 0.f, 0.f}, qh[4] = {0.f, 0.f, 0.f, 0.f};
      if (src_bn) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (bk * 4 + c < K) { qs[c] = __ldg(a.src_scale + bcol + c); qh[c] = __ldg(a.src_shift + bcol + c); }
      }
      for (int r = br; r < tile_rows; r += b_rows) {
        float x[4] = {0.f, 0.f, 0.f, 0.f}, zp[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
          const int64_t row = b0 + r;
          const float* srow = a.src + row * a.ld_src + bcol;
          if (vec_k) {
            const float4 raw = ldg4(srow);
            zp[0] = raw.x; zp[1] = raw.y; zp[2] = raw.z; zp[3] = raw.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (bk * 4 + c < K) zp[c] = __ldg(srow + c);
          }
          if (src_bn) {
            const uint64_t flat = static_cast<uint64_t>(row) * src_width + bcol;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const bool keep = src_thr == 0u || dropout_keep(seed, a.src_salt, flat + c, src_thr);
              x[c] = (bk * 4 + c < K) ? act_value(zp[c], qs[c], qh[c], keep, src_keep_scale) : 0.f;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = zp[c];
          }
        }
        *reinterpret_cast<float4*>(sX + r * ldx + bk * 4) = make_float4(x[0], x[1], x[2], x[3]);
        if (src_bn) *reinterpret_cast<float4*>(sZp + r * ldx + bk * 4) = make_float4(zp[0], zp[1], zp[2], zp[3]);
      }
    }
    __syncthreads();
    if (w_live) {
      for (int r = rs; r < rows; r += rs_n) {
        const float4 d = *reinterpret_cast<const float4*>(sDz + r * ldz + nq * 4);
        const float4 x = *reinterpret_cast<const float4*>(sX + r * ldx + kq * 4);
        aw[0][0] = fmaf(d.x, x.x, aw[0][0]); aw[0][1] = fmaf(d.x, x.y, aw[0][1]);
        aw[0][2] = fmaf(d.x, x.z, aw[0][2]); aw[0][3] = fmaf(d.x, x.w, aw[0][3]);
        aw[1][0] = fmaf(d.y, x.x, aw[1][0]); aw[1][1] = fmaf(d.y, x.y, aw[1][1]);
        aw[1][2] = fmaf(d.y, x.z, aw[1][2]); aw[1][3] = fmaf(d.y, x.w, aw[1][3]);
        aw[2][0] = fmaf(d.z, x.x, aw[2][0]); aw[2][1] = fmaf(d.z, x.y, aw[2][1]);
        aw[2][2] = fmaf(d.z, x.z, aw[2][2]); aw[2][3] = fmaf(d.z, x.w, aw[2][3]);
        aw[3][0] = fmaf(d.w, x.x, aw[3][0]); aw[3][1] = fmaf(d.w, x.y, aw[3][1]);
        aw[3][2] = fmaf(d.w, x.z, aw[3][2]); aw[3][3] = fmaf(d.w, x.w, aw[3][3]);
      }
    }
    if (d_live && ty * 4 < rows) {
      float di[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) di[r][c] = 0.f;
      const float* dz_rows = sDz + (ty * 4) * ldz;
      for (int n0 = 0; n0 < Np; n0 += 4) {
        float4 d4[4], w4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) d4[r] = *reinterpret_cast<const float4*>(dz_rows + r * ldz + n0);
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) w4[nn] = *reinterpret_cast<const float4*>(sW + (n0 + nn) * Kp + c0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float dv[4] = {d4[r].x, d4[r].y, d4[r].z, d4[r].w};
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) {
            di[r][0] = fmaf(dv[nn], w4[nn].x, di[r][0]);
            di[r][1] = fmaf(dv[nn], w4[nn].y, di[r][1]);
            di[r][2] = fmaf(dv[nn], w4[nn].z, di[r][2]);
            di[r][3] = fmaf(dv[nn], w4[nn].w, di[r][3]);
          }
        }
      }
      float pm[4] = {0.f, 0.f, 0.f, 0.f}, pr[4] = {0.f, 0.f, 0.f, 0.f};
      if (src_bn) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c0 + c < K) { pm[c] = __ldg(a.src_mean + g * K + c0 + c); pr[c] = __ldg(a.src_rstd + g * K + c0 + c); }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int lr = ty * 4 + r;
        if (lr >= rows) continue;
        const int64_t row = b0 + lr;
        if (a.d_in) {
          float* dst = a.d_in + row * src_width + g * K + c0;
          if (vec_din) {
            *reinterpret_cast<float4*>(dst) = make_float4(di[r][0], di[r][1], di[r][2], di[r][3]);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c0 + c < K) dst[c] = di[r][c];
          }
        }
        if (src_bn) {
          const float4 xa = *reinterpret_cast<const float4*>(sX + lr * ldx + c0);
          const float4 za = *reinterpret_cast<const float4*>(sZp + lr * ldx + c0);
          const float xv[4] = {xa.x, xa.y, xa.z, xa.w}, zv[4] = {za.x, za.y, za.z, za.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float dy = xv[c] > 0.f ? di[r][c] * src_keep_scale : 0.f;
            s1[c] += dy;
            s2[c] = fmaf(dy, (zv[c] - pm[c]) * pr[c], s2[c]);
          }
        }
      }
    }
  }

  // ---- weight gradient: row subsets in order, then one partial per CTA
  __syncthreads();
  float* sRed = sDz;
  if (w_live) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sRed[(rs * mt_n + mt) * 16 + i * 4 + j] = aw[i][j];
  }
  __syncthreads();
  if (w_live && rs == 0) {
    float* dst = partial_w + (static_cast<int64_t>(blockIdx.x) * G + g) * N * K;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        for (int y = 0; y < rs_n; ++y) v += sRed[(y * mt_n + mt) * 16 + i * 4 + j];
        const int n = nq * 4 + i, k = kq * 4 + j;
        if (n < N && k < K) dst[n * K + k] = v;
      }
  }
  if (!src_bn) return;
  // ---- BatchNorm-backward sums of the layer below
  __syncthreads();
  if (d_live) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sRed[(ty * 2 + 0) * Kp + c0 + c] = s1[c];
      sRed[(ty * 2 + 1) * Kp + c0 + c] = s2[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float u = 0.f, v = 0.f;
    for (int y = 0; y < ty_n; ++y) {
      u += sRed[(y * 2 + 0) * Kp + threadIdx.x];
      v += sRed[(y * 2 + 1) * Kp + threadIdx.x];
    }
    partial_s[(static_cast<int64_t>(blockIdx.x) * 2 + 0) * src_width + g * K + threadIdx.x] = u;
    partial_s[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * src_width + g * K + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kThreads) hei_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial,
                                                                    int64_t n, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  for (int p = 0; p < n_partial; ++p) v += partial[static_cast<int64_t>(p) * n + i];
  out[i] = v;
}

struct Tiling {
  int tx_n, ty_n, tile_rows, tiles_per_cta, n_cta;
};
// quads = output columns / 4 of the per-row GEMM; ctas_per_group = how many row chunks a group is cut into
Tiling tiling(int64_t m, int quads, int groups, int sm_multiple) {
  Tiling t;
  t.tx_n = quads;
  t.ty_n = kThreads / quads;
  if (t.ty_n > kMaxTileQuads) t.ty_n = kMaxTileQuads;
  t.tile_rows = t.ty_n * 4;
  const int64_t n_tiles = (m + t.tile_rows - 1) / t.tile_rows;
  int64_t cap = static_cast<int64_t>(kNumSMs) * sm_multiple / groups;
  if (cap < 1) cap = 1;
  if (cap > kStatCtas) cap = kStatCtas;
  t.tiles_per_cta = static_cast<int>((n_tiles + cap - 1) / cap);
  if (t.tiles_per_cta < 1) t.tiles_per_cta = 1;
  t.n_cta = static_cast<int>((n_tiles + t.tiles_per_cta - 1) / t.tiles_per_cta);
  if (t.n_cta < 1) t.n_cta = 1;
  return t;
}

// Upper bounds on the CTAs per SM that can be resident at once (registers); the launch asks the runtime for the
// real figure at its shared-memory size and makes the grid ONE full wave of them.
constexpr int kFwdSmMultiple = 3, kBwdSmMultiple = 3;

template <typename Kernel>
int resident_ctas(Kernel kernel, size_t smem, int fallback) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem) != cudaSuccess || n < 1) n = fallback;
  return n;
}

size_t fwd_smem(const Tiling& t, int K, int N) {
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  const size_t tile = static_cast<size_t>(t.tile_rows) * (Kp + 4), red = static_cast<size_t>(t.ty_n) * 2 * Np;
  return sizeof(float) * (static_cast<size_t>(Kp) * Np + Np + (tile > red ? tile : red));
}
size_t bwd_smem(const Tiling& t, int K, int N, bool src_bn) {
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  size_t tile = static_cast<size_t>(t.tile_rows) * ((Np + 4) + (Kp + 4) * (src_bn ? 2 : 1));
  const size_t red_w = static_cast<size_t>(kThreads) * 16, red_s = static_cast<size_t>(t.ty_n) * 2 * Kp;
  if (tile < red_w) tile = red_w;
  if (tile < red_s) tile = red_s;
  return sizeof(float) * (static_cast<size_t>(Np) * Kp + tile);
}

bool shape_ok(int groups, int k, int n) {
  return groups > 0 && groups <= 65535 && k > 0 && n > 0 && k <= kMaxWidth && n <= kMaxWidth;
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_hei_layer_supported(int32_t groups, int32_t k, int32_t n) { return aread::shape_ok(groups, k, n) ? 1 : 0; }

void aread_hei_set_path(int32_t tensor_cores_fwd, int32_t tensor_cores_bwd) {
  aread::hei_tc_set_path(tensor_cores_fwd, tensor_cores_bwd);
}

int aread_hei_layer_path(int64_t m, int32_t groups, int32_t k, int32_t n) {
  alignas(16) static const float aligned = 0.f;     // a 16-byte aligned, stride-4 source: only the shape decides
  return (aread::hei_tc_usable(m, groups, k, n, &aligned, 4) ? 1 : 0) |
         (aread::hei_tc_bwd_usable(m, groups, k, n, &aligned, 4) ? 2 : 0);
}

size_t aread_hei_layer_workspace_bytes(int64_t m, int32_t groups, int32_t k, int32_t n) {
  using namespace aread;
  if (!shape_ok(groups, k, n) || m <= 0) return 256;
  const Tiling tf = tiling(m, ((n + 3) & ~3) / 4, groups, kFwdSmMultiple);
  const Tiling tb = tiling(m, ((k + 3) & ~3) / 4, groups, kBwdSmMultiple);
  const size_t fwd = static_cast<size_t>(tf.n_cta) * 2 * groups * n;
  const size_t bwd = static_cast<size_t>(tb.n_cta) * groups * n * k + static_cast<size_t>(tb.n_cta) * 2 * groups * k;
  size_t need = fwd > bwd ? fwd : bwd;
  const size_t tc = hei_tc_workspace_floats(m, groups, k, n);       // the tensor-core path (hei_tc.cu)
  if (tc > need) need = tc;
  return align_up(sizeof(float) * need, 256);
}

int aread_hei_layer_fwd(const aread_hei_layer_fwd_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "hei_layer_fwd: null args");
  const aread_hei_layer_fwd_args& a = *args;
  AREAD_REQUIRE(a.m > 0 && shape_ok(a.groups, a.k, a.n), "hei_layer_fwd: unsupported shape m=%lld groups=%d k=%d n=%d",
                (long long)a.m, a.groups, a.k, a.n);
  AREAD_REQUIRE(a.src && a.weight && a.z && a.mean && a.rstd && a.scale && a.shift, "hei_layer_fwd: null pointer");
  AREAD_REQUIRE((a.src_scale == nullptr) == (a.src_shift == nullptr), "hei_layer_fwd: src_scale / src_shift mismatch");
  AREAD_REQUIRE(a.bn_skip || (a.gamma && a.beta && a.running_mean && a.running_var), "hei_layer_fwd: null BN tensor");
  AREAD_REQUIRE(a.src_p >= 0.f && a.src_p < 1.f, "hei_layer_fwd: dropout %f not in [0, 1)", a.src_p);
  AREAD_REQUIRE(a.ld_src >= static_cast<int64_t>(a.groups) * a.k, "hei_layer_fwd: ld_src too small");
  AREAD_REQUIRE(a.workspace && a.workspace_bytes >= aread_hei_layer_workspace_bytes(a.m, a.groups, a.k, a.n),
                "hei_layer_fwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (hei_tc_usable(a.m, a.groups, a.k, a.n, a.src, a.ld_src)) return hei_tc_fwd(a, stream);
  const int Np = (a.n + 3) & ~3;
  Tiling t = tiling(a.m, Np / 4, a.groups, kFwdSmMultiple);
  const size_t smem = fwd_smem(t, a.k, a.n);
  static uint64_t configured = 0;          // cudaFuncSetAttribute is per device
  if (first_use_on_device(&configured)) {
    AREAD_CUDA(cudaFuncSetAttribute(hei_layer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  AREAD_REQUIRE(smem <= 160 * 1024, "hei_layer_fwd: tile needs %zu bytes of shared memory", smem);
  {  // fewer, longer CTAs when fewer than kFwdSmMultiple fit per SM (never more: the workspace is sized for that)
    const int occ = resident_ctas(hei_layer_fwd_kernel, smem, 1);
    if (occ < kFwdSmMultiple) t = tiling(a.m, Np / 4, a.groups, occ);
  }
  const Drop d = make_drop(a.training ? a.src_p : 0.f);
  const int do_stats = (a.training && !a.bn_skip) ? 1 : 0;
  float* partial = static_cast<float*>(a.workspace);
  AREAD_LAUNCH(hei_layer_fwd_kernel, dim3(t.n_cta, a.groups), kThreads, smem, stream, a, t.tx_n, t.ty_n, t.tiles_per_cta,
               d.threshold, d.keep_scale, do_stats, partial);
  aread_bn_act_args f = {};
  f.m = a.m;
  f.width = a.groups * a.n;
  f.training = a.training;
  f.bn_skip = a.bn_skip;
  f.momentum = a.momentum;
  f.eps = a.eps;
  f.z = a.z;
  f.ldz = f.width;
  f.gamma = a.gamma;
  f.beta = a.beta;
  f.running_mean = a.running_mean;
  f.running_var = a.running_var;
  f.mean = a.mean;
  f.rstd = a.rstd;
  f.scale = a.scale;
  f.shift = a.shift;
  AREAD_LAUNCH(bn_finalize_kernel, ceil_div(f.width, 32), kBnThreads, 0, stream, f, partial, t.n_cta);
  return AREAD_OK;
}

int aread_hei_layer_bwd(const aread_hei_layer_bwd_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "hei_layer_bwd: null args");
  const aread_hei_layer_bwd_args& a = *args;
  AREAD_REQUIRE(a.m > 0 && shape_ok(a.groups, a.k, a.n), "hei_layer_bwd: unsupported shape m=%lld groups=%d k=%d n=%d",
                (long long)a.m, a.groups, a.k, a.n);
  AREAD_REQUIRE(a.z && a.d_out && a.mean && a.rstd && a.scale && a.shift && a.coef && a.src && a.weight && a.d_w,
                "hei_layer_bwd: null pointer");
  const bool src_bn = a.src_scale != nullptr;
  AREAD_REQUIRE(!src_bn || (a.src_shift && a.src_mean && a.src_rstd && a.src_coef), "hei_layer_bwd: incomplete src BN");
  AREAD_REQUIRE(a.p >= 0.f && a.p < 1.f && a.src_p >= 0.f && a.src_p < 1.f, "hei_layer_bwd: bad dropout");
  AREAD_REQUIRE(a.ld_src >= static_cast<int64_t>(a.groups) * a.k, "hei_layer_bwd: ld_src too small");
  AREAD_REQUIRE(a.workspace && a.workspace_bytes >= aread_hei_layer_workspace_bytes(a.m, a.groups, a.k, a.n),
                "hei_layer_bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (hei_tc_bwd_usable(a.m, a.groups, a.k, a.n, a.src, a.ld_src)) return hei_tc_bwd(a, stream);
  const int Np = (a.n + 3) & ~3, Kp = (a.k + 3) & ~3;
  Tiling t = tiling(a.m, Kp / 4, a.groups, kBwdSmMultiple);
  const int mt_n = (Np / 4) * (Kp / 4);
  const int rs_n = kThreads / mt_n;
  const size_t smem = bwd_smem(t, a.k, a.n, src_bn);
  static uint64_t configured = 0;          // cudaFuncSetAttribute is per device
  if (first_use_on_device(&configured)) {
    AREAD_CUDA(cudaFuncSetAttribute(hei_layer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  AREAD_REQUIRE(smem <= 200 * 1024, "hei_layer_bwd: tile needs %zu bytes of shared memory", smem);
  {
    const int occ = resident_ctas(hei_layer_bwd_kernel, smem, 1);
    if (occ < kBwdSmMultiple) t = tiling(a.m, Kp / 4, a.groups, occ);
  }
  const Drop d = make_drop(a.p), ds = make_drop(a.src_p);
  float* partial_w = static_cast<float*>(a.workspace);
  float* partial_s = partial_w + static_cast<size_t>(t.n_cta) * a.groups * a.n * a.k;
  AREAD_LAUNCH(hei_layer_bwd_kernel, dim3(t.n_cta, a.groups), kThreads, smem, stream, a, t.tx_n, t.ty_n, mt_n, rs_n,
               t.tiles_per_cta, d.threshold, d.keep_scale, ds.threshold, ds.keep_scale, partial_w, partial_s);
  const int64_t n_w = static_cast<int64_t>(a.groups) * a.n * a.k;
  AREAD_LAUNCH(hei_wgrad_reduce_kernel, ceil_div(n_w, kThreads), kThreads, 0, stream, partial_w, t.n_cta, n_w, a.d_w);
  if (src_bn) {
    aread_bn_act_bwd_args f = {};
    f.m = a.m;
    f.width = a.groups * a.k;
    f.bn_skip = a.bn_skip;
    f.d_gamma = a.src_d_gamma;
    f.d_beta = a.src_d_beta;
    f.d_bias = a.src_d_bias;
    AREAD_LAUNCH(bn_bwd_finalize_kernel, ceil_div(f.width, 32), kBnThreads, 0, stream, f, partial_s, t.n_cta, a.src_coef);
  }
  return AREAD_OK;
}

}  // extern "C"
