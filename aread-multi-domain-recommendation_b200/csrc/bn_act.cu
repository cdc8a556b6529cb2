// BatchNorm1d + ReLU + Dropout around the grouped Linear layers, forward and backward, and the
// MMoE gate mixture.  All of it is HBM-bound elementwise / column-reduction work in fp32.
//
// Reference arithmetic: MultiLayerPerceptron.forward (model/layer.py:221-229): Linear ->
// BatchNorm1d (training: batch mean / biased variance, eps 1e-5, running stats with momentum 0.1
// and the unbiased variance; eval: running stats; skipped for a batch of one, layer.py:226) ->
// ReLU -> Dropout(p); MMoE mixture: model/aread.py:150-153.
//
// Column reductions are deterministic: every CTA reduces a fixed row range into a partial, and the
// partials are summed in CTA order.
#include "common.cuh"
#include "bn_common.cuh"

namespace aread {
namespace {

constexpr int kThreads = kBnThreads;

struct Geometry {  // how 256 threads cover [rows, width] in the column-sum kernels
  int tw, vec;
};
inline Geometry geometry(int width, bool aligned) {
  Geometry g;
  g.vec = (aligned && width % 4 == 0) ? 4 : 1;
  g.tw = 1;
  while (g.tw * g.vec < width && g.tw < kThreads) g.tw <<= 1;
  return g;
}

// Per-CTA partial column sums of the two quantities f produces, over the CTA's contiguous row range.
// Threads: tx over column groups of VEC, ty over rows; four independent row loads are in flight per
// thread.  The ty partials are added in ty order, the CTA partials later in CTA order.
// `prep(col)` runs once per column group and its result is handed to every f call of that group: per-column
// parameters live in registers instead of being reloaded for every row.
template <int VEC, typename P, typename F>
__device__ __forceinline__ void column_partials(int64_t m, int width, int tw, float* partial, P prep, F f) {
  extern __shared__ float s_red[];  // [ty_n][2][tw * VEC]
  const int ty_n = kThreads / tw;
  const int tx = threadIdx.x % tw, ty = threadIdx.x / tw;
  const int64_t rows_per_cta = (m + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per_cta;
  const int64_t r1 = min(m, r0 + rows_per_cta);
  const int span = tw * VEC;
  for (int c0 = 0; c0 < width; c0 += span) {
    const int col = c0 + tx * VEC;
    float s[VEC], q[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = q[v] = 0.f;
    if (col < width) {
      const auto prm = prep(col);
      int64_t r = r0 + ty;
      for (; r + 3 * ty_n < r1; r += 4 * ty_n) {
        float a[4][VEC], b[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) f(r + u * ty_n, col, prm, a[u], b[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < VEC; ++v) { s[v] += a[u][v]; q[v] += b[u][v]; }
      }
      for (; r < r1; r += ty_n) {
        float a[VEC], b[VEC];
        f(r, col, prm, a, b);
#pragma unroll
        for (int v = 0; v < VEC; ++v) { s[v] += a[v]; q[v] += b[v]; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      s_red[(ty * 2 + 0) * span + tx * VEC + v] = s[v];
      s_red[(ty * 2 + 1) * span + tx * VEC + v] = q[v];
    }
    __syncthreads();
    if (ty == 0 && col < width) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float a = 0.f, b = 0.f;
        for (int y = 0; y < ty_n; ++y) {
          a += s_red[(y * 2 + 0) * span + tx * VEC + v];
          b += s_red[(y * 2 + 1) * span + tx * VEC + v];
        }
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + 0) * width + col + v] = a;
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * width + col + v] = b;
      }
    }
  }
}

// keep flags of VEC consecutive elements starting at the (even, when VEC == 4) flat index: one hash per pair
template <int VEC>
__device__ __forceinline__ void keep_flags(uint64_t seed, uint32_t salt, uint64_t flat, uint32_t threshold,
                                           bool (&keep)[VEC]) {
  if (threshold == 0u) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) keep[v] = true;
  } else if (VEC == 4 && (flat & 1) == 0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t h = dropout_pair_hash(seed, salt, (flat >> 1) + q);
      keep[2 * q] = (h & 0xffffu) >= threshold;
      keep[2 * q + 1] = (h >> 16) >= threshold;
    }
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) keep[v] = dropout_keep(seed, salt, flat + v, threshold);
  }
}

// ---------------------------------------------------------------------------------- forward
template <int VEC>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(int64_t m, int width, int tw, const float* __restrict__ z,
                                                            int64_t ldz, float* __restrict__ partial) {
  // sums are taken about the column's first row (a sample of the same distribution), which keeps
  // E[v^2] - E[v]^2 well conditioned whatever the column mean is
  struct Pivot {
    float p[VEC];
  };
  column_partials<VEC>(
      m, width, tw, partial,
      [&](int col) {
        Pivot pv;
#pragma unroll
        for (int v = 0; v < VEC; ++v) pv.p[v] = z[col + v];
        return pv;
      },
      [&](int64_t r, int col, const Pivot& pv, float (&a)[VEC], float (&b)[VEC]) {
        if (VEC == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(z + r * ldz + col));
          a[0] = v.x - pv.p[0]; a[1] = v.y - pv.p[1]; a[2] = v.z - pv.p[2]; a[3] = v.w - pv.p[3];
        } else {
          a[0] = __ldg(z + r * ldz + col) - pv.p[0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) b[v] = a[v] * a[v];
      });
}

// rows by (blockIdx.x, ty), columns by tx, VEC consecutive columns per thread: no index divisions
template <int VEC>
__global__ void __launch_bounds__(kThreads) bn_act_kernel(const aread_bn_act_args a, int tw, uint32_t threshold,
                                                          float keep_scale) {
  const uint64_t seed = seed_of(a);
  const int tx = threadIdx.x % tw, ty = threadIdx.x / tw, ty_n = kThreads / tw;
  for (int col = tx * VEC; col < a.width; col += tw * VEC) {
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sc[j] = a.scale[col + j];
      sh[j] = a.shift[col + j];
    }
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * ty_n + ty; r < a.m; r += static_cast<int64_t>(gridDim.x) * ty_n) {
      float z[VEC], v[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.z + r * a.ldz + col));
        z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
      } else {
        z[0] = __ldg(a.z + r * a.ldz + col);
      }
      bool keep[VEC];
      keep_flags<VEC>(seed, a.salt, static_cast<uint64_t>(r) * a.width + col, threshold, keep);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = act_value(z[j], sc[j], sh[j], keep[j], keep_scale);
      if (VEC == 4) {
        if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + r * a.ldo + col) = make_float4(v[0], v[1], v[2], v[3]);
        if (a.out_bf16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<unsigned*>(&lo);
          pk.y = *reinterpret_cast<unsigned*>(&hi);
          *reinterpret_cast<uint2*>(a.out_bf16 + r * a.ldo + col) = pk;
          if (a.out_bf16_lo) {
            const float2 f0 = __bfloat1622float2(lo), f1 = __bfloat1622float2(hi);
            __nv_bfloat162 rl = __floats2bfloat162_rn(v[0] - f0.x, v[1] - f0.y);
            __nv_bfloat162 rh = __floats2bfloat162_rn(v[2] - f1.x, v[3] - f1.y);
            pk.x = *reinterpret_cast<unsigned*>(&rl);
            pk.y = *reinterpret_cast<unsigned*>(&rh);
            *reinterpret_cast<uint2*>(a.out_bf16_lo + r * a.ldo + col) = pk;
          }
        }
      } else {
        if (a.out_f32) a.out_f32[r * a.ldo + col] = v[0];
        if (a.out_bf16) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[0]);
          reinterpret_cast<__nv_bfloat16*>(a.out_bf16)[r * a.ldo + col] = h;
          if (a.out_bf16_lo)
            reinterpret_cast<__nv_bfloat16*>(a.out_bf16_lo)[r * a.ldo + col] =
                __float2bfloat16_rn(v[0] - __bfloat162float(h));
        }
      }
    }
  }
}

// the pre-activation of row r, columns col .. col + VEC - 1, from the fp32 or the bf16 copy
template <int VEC, typename Args>
__device__ __forceinline__ void load_z(const Args& a, int64_t r, int col, float (&z)[VEC]) {
  if (a.z_bf16 != nullptr) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.z_bf16) + r * a.ldz + col;
    if (VEC == 4) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
      z[0] = __uint_as_float(w.x << 16);
      z[1] = __uint_as_float(w.x & 0xffff0000u);
      z[2] = __uint_as_float(w.y << 16);
      z[3] = __uint_as_float(w.y & 0xffff0000u);
    } else {
      z[0] = __bfloat162float(*p);
    }
  } else if (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.z + r * a.ldz + col));
    z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
  } else {
    z[0] = __ldg(a.z + r * a.ldz + col);
  }
}
template <typename Args>
__device__ __forceinline__ float load_z1(const Args& a, int64_t idx) {
  return a.z_bf16 != nullptr ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.z_bf16)[idx]) : __ldg(a.z + idx);
}

// ---------------------------------------------------------------------------------- backward
// dy = d_out * [y > 0] * keep / (1 - p);  xhat = (z - mean) * rstd
template <int VEC>
__global__ void __launch_bounds__(kThreads) bn_bwd_stats_kernel(const aread_bn_act_bwd_args a, int tw,
                                                                uint32_t threshold, float keep_scale,
                                                                float* __restrict__ partial) {
  const uint64_t seed = seed_of(a);
  struct Params {
    float sc[VEC], sh[VEC], rs[VEC], mr[VEC];    // scale, shift, rstd, mean * rstd
  };
  column_partials<VEC>(
      a.m, a.width, tw, partial,
      [&](int col) {
        Params pr;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          pr.sc[v] = a.scale[col + v];
          pr.sh[v] = a.shift[col + v];
          pr.rs[v] = a.rstd[col + v];
          pr.mr[v] = a.mean[col + v] * pr.rs[v];
        }
        return pr;
      },
      [&](int64_t r, int col, const Params& pr, float (&s1)[VEC], float (&s2)[VEC]) {
        float z[VEC], d[VEC];
        load_z<VEC>(a, r, col, z);
        if (VEC == 4) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(a.d_out + r * a.ldd + col));
          d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
        } else {
          d[0] = __ldg(a.d_out + r * a.ldd + col);
        }
        bool keep[VEC];
        keep_flags<VEC>(seed, a.salt, static_cast<uint64_t>(r) * a.width + col, threshold, keep);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float y = fmaf(z[v], pr.sc[v], pr.sh[v]);
          const float dy = (y > 0.f && keep[v]) ? d[v] * keep_scale : 0.f;
          s1[v] = dy;
          s2[v] = dy * fmaf(z[v], pr.rs[v], -pr.mr[v]);
        }
      });
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const aread_bn_act_bwd_args a, int tw,
                                                                uint32_t threshold, float keep_scale,
                                                                const float* __restrict__ coef) {
  const uint64_t seed = seed_of(a);
  const int tx = threadIdx.x % tw, ty = threadIdx.x / tw, ty_n = kThreads / tw;
  for (int col = tx * VEC; col < a.width; col += tw * VEC) {
    float sc[VEC], sh[VEC], rs[VEC], mr[VEC], c0[VEC], c1[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sc[j] = a.scale[col + j];
      sh[j] = a.shift[col + j];
      rs[j] = a.rstd[col + j];
      mr[j] = a.mean[col + j] * rs[j];
      c0[j] = coef[col + j];
      c1[j] = coef[a.width + col + j];
    }
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * ty_n + ty; r < a.m; r += static_cast<int64_t>(gridDim.x) * ty_n) {
      float z[VEC], d[VEC], dz[VEC];
      load_z<VEC>(a, r, col, z);
      if (VEC == 4) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(a.d_out + r * a.ldd + col));
        d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
      } else {
        d[0] = __ldg(a.d_out + r * a.ldd + col);
      }
      bool keep[VEC];
      keep_flags<VEC>(seed, a.salt, static_cast<uint64_t>(r) * a.width + col, threshold, keep);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float y = fmaf(z[j], sc[j], sh[j]);
        const float dy = (y > 0.f && keep[j]) ? d[j] * keep_scale : 0.f;
        const float xhat = fmaf(z[j], rs[j], -mr[j]);
        dz[j] = a.bn_skip ? dy : sc[j] * (dy - c0[j] - xhat * c1[j]);
      }
      if (VEC == 4) {
        if (a.dz_f32) *reinterpret_cast<float4*>(a.dz_f32 + r * a.ldo + col) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        if (a.dz_bf16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(dz[0], dz[1]), hi = __floats2bfloat162_rn(dz[2], dz[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<unsigned*>(&lo);
          pk.y = *reinterpret_cast<unsigned*>(&hi);
          *reinterpret_cast<uint2*>(a.dz_bf16 + r * a.ldo + col) = pk;
          if (a.dz_bf16_lo) {
            const float2 f0 = __bfloat1622float2(lo), f1 = __bfloat1622float2(hi);
            __nv_bfloat162 rl = __floats2bfloat162_rn(dz[0] - f0.x, dz[1] - f0.y);
            __nv_bfloat162 rh = __floats2bfloat162_rn(dz[2] - f1.x, dz[3] - f1.y);
            pk.x = *reinterpret_cast<unsigned*>(&rl);
            pk.y = *reinterpret_cast<unsigned*>(&rh);
            *reinterpret_cast<uint2*>(a.dz_bf16_lo + r * a.ldo + col) = pk;
          }
        }
      } else {
        if (a.dz_f32) a.dz_f32[r * a.ldo + col] = dz[0];
        if (a.dz_bf16) {
          const __nv_bfloat16 h = __float2bfloat16_rn(dz[0]);
          reinterpret_cast<__nv_bfloat16*>(a.dz_bf16)[r * a.ldo + col] = h;
          if (a.dz_bf16_lo)
            reinterpret_cast<__nv_bfloat16*>(a.dz_bf16_lo)[r * a.ldo + col] =
                __float2bfloat16_rn(dz[0] - __bfloat162float(h));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- MMoE mixture
// h[b, e, :] = dropout(relu(bn(z[b, e, :]))) ;  out[b, g, :] = sum_e gate[b, g, e] * h[b, e, :]
// NE / NG are compile-time bounds of the loops (NE_ == 0: generic, bounded by 16 / 8) so that the per-thread
// arrays live in registers.
template <int NE_, int NG_>
__global__ void __launch_bounds__(kThreads) mmoe_mix_fwd_kernel(const aread_mmoe_mix_args a, uint32_t threshold,
                                                                float keep_scale) {
  const uint64_t seed = seed_of(a);
  constexpr int MAXE = NE_ > 0 ? NE_ : 16;
  const int H = a.width, NE = NE_ > 0 ? NE_ : a.n_expert, G = NG_ > 0 ? NG_ : a.n_gate;
  const int64_t total = a.m * H;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / H;
    const int c = static_cast<int>(i - b * H);
    float h[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      if (e < NE) {
        const int col = e * H + c;
        const bool keep = threshold == 0u ||
                          dropout_keep(seed, a.salt, static_cast<uint64_t>(b) * (NE * H) + col, threshold);
        h[e] = act_value(load_z1(a, b * a.ldz + col), __ldg(a.scale + col), __ldg(a.shift + col), keep, keep_scale);
      }
    }
    for (int g = 0; g < G; ++g) {
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e)
        if (e < NE) acc = fmaf(__ldg(a.gate + b * (G * NE) + g * NE + e), h[e], acc);
      a.out[b * (G * H) + g * H + c] = acc;
    }
  }
}

// Four consecutive columns per thread (128-bit / 64-bit loads of z, one hash per element pair, 128-bit stores):
// H % 4 == 0, ldz % 4 == 0, 16-byte aligned tensors.
template <int NE_, int NG_>
__global__ void __launch_bounds__(kThreads) mmoe_mix_fwd_vec4_kernel(const aread_mmoe_mix_args a, uint32_t threshold,
                                                                     float keep_scale) {
  const uint64_t seed = seed_of(a);
  constexpr int MAXE = NE_ > 0 ? NE_ : 16;
  const int H = a.width, H4 = a.width / 4, NE = NE_ > 0 ? NE_ : a.n_expert, G = NG_ > 0 ? NG_ : a.n_gate;
  const int64_t total = a.m * H4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / H4;
    const int c = static_cast<int>(i - b * H4) * 4;
    float h[MAXE][4];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      if (e < NE) {
        const int col = e * H + c;
        float z[4];
        load_z<4>(a, b, col, z);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + col));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.shift + col));
        bool keep[4];
        keep_flags<4>(seed, a.salt, static_cast<uint64_t>(b) * (NE * H) + col, threshold, keep);
        h[e][0] = act_value(z[0], sc.x, sh.x, keep[0], keep_scale);
        h[e][1] = act_value(z[1], sc.y, sh.y, keep[1], keep_scale);
        h[e][2] = act_value(z[2], sc.z, sh.z, keep[2], keep_scale);
        h[e][3] = act_value(z[3], sc.w, sh.w, keep[3], keep_scale);
      }
    }
    for (int g = 0; g < G; ++g) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int e = 0; e < MAXE; ++e)
        if (e < NE) {
          const float w = __ldg(a.gate + b * (G * NE) + g * NE + e);
          acc.x = fmaf(w, h[e][0], acc.x); acc.y = fmaf(w, h[e][1], acc.y);
          acc.z = fmaf(w, h[e][2], acc.z); acc.w = fmaf(w, h[e][3], acc.w);
        }
      *reinterpret_cast<float4*>(a.out + b * (G * H) + g * H + c) = acc;
    }
  }
}

// d_h[b, e, :] = sum_g gate[b, g, e] * d_out[b, g, :] ;  d_gate[b, g, e] = <d_out[b, g, :], h[b, e, :]>
// one warp per sample
template <int NE_, int NG_>
__global__ void __launch_bounds__(kThreads) mmoe_mix_bwd_kernel(const aread_mmoe_mix_args a, uint32_t threshold,
                                                                float keep_scale) {
  const uint64_t seed = seed_of(a);
  constexpr int MAXE = NE_ > 0 ? NE_ : 16, MAXG = NG_ > 0 ? NG_ : 8;
  const int H = a.width, NE = NE_ > 0 ? NE_ : a.n_expert, G = NG_ > 0 ? NG_ : a.n_gate;
  const int lane = threadIdx.x % 32;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x / 32);
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x / 32) + threadIdx.x / 32; b < a.m; b += warps) {
    float dg[MAXG][MAXE], gt[MAXG][MAXE];
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        dg[g][e] = 0.f;
        gt[g][e] = (g < G && e < NE) ? __ldg(a.gate + b * (G * NE) + g * NE + e) : 0.f;
      }
    for (int c = lane; c < H; c += 32) {
      float dout[MAXG];
#pragma unroll
      for (int g = 0; g < MAXG; ++g) dout[g] = g < G ? __ldg(a.d_out + b * (G * H) + g * H + c) : 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < NE) {
          const int col = e * H + c;
          const bool keep = threshold == 0u ||
                            dropout_keep(seed, a.salt, static_cast<uint64_t>(b) * (NE * H) + col, threshold);
          const float h = act_value(load_z1(a, b * a.ldz + col), __ldg(a.scale + col), __ldg(a.shift + col), keep,
                                    keep_scale);
          float acc = 0.f;
#pragma unroll
          for (int g = 0; g < MAXG; ++g) {
            acc = fmaf(gt[g][e], dout[g], acc);
            dg[g][e] = fmaf(dout[g], h, dg[g][e]);
          }
          a.d_h[b * (NE * H) + col] = acc;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
      for (int e = 0; e < MAXE; ++e)
        if (g < G && e < NE) {
          const float v = warp_sum(dg[g][e]);
          if (lane == 0) a.d_gate[b * (G * NE) + g * NE + e] = v;
        }
  }
}

// Four consecutive columns per lane, H / 4 lanes per sample (H in {16, 32, 64, 128}: 8 .. 1 samples per warp): 64-bit /
// 128-bit loads, one hash per element pair, 128-bit stores, and the gate-gradient dot products reduced over the lanes of
// a sample by a log2(H / 4)-step butterfly.
template <int NE_, int NG_>
__global__ void __launch_bounds__(kThreads) mmoe_mix_bwd_vec4_kernel(const aread_mmoe_mix_args a, uint32_t threshold,
                                                                     float keep_scale) {
  const uint64_t seed = seed_of(a);
  constexpr int MAXE = NE_ > 0 ? NE_ : 16, MAXG = NG_ > 0 ? NG_ : 8;
  const int H = a.width, NE = NE_ > 0 ? NE_ : a.n_expert, G = NG_ > 0 ? NG_ : a.n_gate;
  const int lpr = H / 4, rpw = 32 / lpr;                      // lanes per sample, samples per warp
  const int lane = threadIdx.x % 32;
  const int sub = lane / lpr, c = (lane - sub * lpr) * 4;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x / 32);
  const int64_t n_groups = (a.m + rpw - 1) / rpw;
  for (int64_t grp = static_cast<int64_t>(blockIdx.x) * (blockDim.x / 32) + threadIdx.x / 32; grp < n_groups; grp += warps) {
    const int64_t b = grp * rpw + sub;
    const bool valid = b < a.m;
    const int64_t bb = valid ? b : 0;
    float dg[MAXG][MAXE], gt[MAXG][MAXE];
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        dg[g][e] = 0.f;
        gt[g][e] = (g < G && e < NE) ? __ldg(a.gate + bb * (G * NE) + g * NE + e) : 0.f;
      }
    float4 dout[MAXG];
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
      dout[g] = g < G ? __ldg(reinterpret_cast<const float4*>(a.d_out + bb * (G * H) + g * H + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      if (e < NE) {
        const int col = e * H + c;
        float z[4];
        load_z<4>(a, bb, col, z);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + col));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.shift + col));
        bool keep[4];
        keep_flags<4>(seed, a.salt, static_cast<uint64_t>(bb) * (NE * H) + col, threshold, keep);
        const float h0 = act_value(z[0], sc.x, sh.x, keep[0], keep_scale), h1 = act_value(z[1], sc.y, sh.y, keep[1], keep_scale);
        const float h2 = act_value(z[2], sc.z, sh.z, keep[2], keep_scale), h3 = act_value(z[3], sc.w, sh.w, keep[3], keep_scale);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < MAXG; ++g) {
          acc.x = fmaf(gt[g][e], dout[g].x, acc.x); acc.y = fmaf(gt[g][e], dout[g].y, acc.y);
          acc.z = fmaf(gt[g][e], dout[g].z, acc.z); acc.w = fmaf(gt[g][e], dout[g].w, acc.w);
          dg[g][e] = fmaf(dout[g].x, h0, fmaf(dout[g].y, h1, fmaf(dout[g].z, h2, fmaf(dout[g].w, h3, dg[g][e]))));
        }
        if (valid) *reinterpret_cast<float4*>(a.d_h + b * (NE * H) + col) = acc;
      }
    }
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
#pragma unroll
      for (int e = 0; e < MAXE; ++e)
        if (g < G && e < NE) {
          float v = dg[g][e];
          for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (valid && lane == sub * lpr) a.d_gate[b * (G * NE) + g * NE + e] = v;
        }
  }
}

__global__ void __launch_bounds__(kThreads) dropout_mask_kernel(uint64_t seed, uint32_t salt, int64_t n,
                                                                uint32_t threshold, uint8_t* out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (threshold == 0u || dropout_keep(seed, salt, static_cast<uint64_t>(i), threshold)) ? 1 : 0;
}

struct RowGrid {
  int tw, vec;
  unsigned grid;
};
// thread layout of the row-major elementwise kernels: vec = 4 when every row start is 16-byte aligned
RowGrid row_grid(int64_t m, int width, bool aligned) {
  RowGrid g;
  g.vec = (aligned && width % 4 == 0) ? 4 : 1;
  g.tw = 1;
  while (g.tw * g.vec < width && g.tw < kThreads) g.tw <<= 1;
  const int rows_per_cta = kThreads / g.tw;
  int64_t n = (m + rows_per_cta - 1) / rows_per_cta;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  g.grid = static_cast<unsigned>(n < 1 ? 1 : (n > cap ? cap : n));
  return g;
}

unsigned elementwise_grid(int64_t total) {
  int64_t g = (total + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

int stat_ctas(int64_t m) {
  int64_t c = (m + 255) / 256;  // at least 256 rows per CTA
  return static_cast<int>(c < 1 ? 1 : (c > kStatCtas ? kStatCtas : c));
}

}  // namespace
}  // namespace aread

extern "C" {

size_t aread_bn_workspace_bytes(int32_t width) {
  return aread::align_up(static_cast<size_t>(aread::kStatCtas) * 2 * width * 4 + 2 * static_cast<size_t>(width) * 4,
                         256);
}

int aread_bn_act_fwd(const aread_bn_act_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "bn_act_fwd: null args");
  const aread_bn_act_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.width > 0 && a.width <= 65536, "bn_act_fwd: width %d unsupported",
                a.width);
  AREAD_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, "bn_act_fwd: dropout %f not in [0, 1)", a.dropout_p);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.z && a.mean && a.rstd && a.scale && a.shift, "bn_act_fwd: null pointer");
  AREAD_REQUIRE(a.bn_skip || (a.gamma && a.beta && a.running_mean && a.running_var), "bn_act_fwd: null BN tensor");
  AREAD_REQUIRE(a.workspace_bytes >= aread_bn_workspace_bytes(a.width), "bn_act_fwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partial = static_cast<float*>(a.workspace);
  const int n_partial = stat_ctas(a.m);
  if (a.training && !a.bn_skip) {
    const Geometry g = geometry(a.width, a.ldz % 4 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0);
    const size_t smem = sizeof(float) * 2 * kThreads * g.vec;
    if (g.vec == 4)
      AREAD_LAUNCH(bn_stats_kernel<4>, n_partial, kThreads, smem, stream, a.m, a.width, g.tw, a.z, a.ldz, partial);
    else
      AREAD_LAUNCH(bn_stats_kernel<1>, n_partial, kThreads, smem, stream, a.m, a.width, g.tw, a.z, a.ldz, partial);
  }
  AREAD_LAUNCH(bn_finalize_kernel, ceil_div(a.width, 32), kThreads, 0, stream, a, partial, n_partial);
  if (a.out_f32 || a.out_bf16) {
    const bool drop = a.training && a.dropout_p > 0.f;
    const uint32_t threshold = drop ? dropout_threshold(a.dropout_p) : 0u;
    const float keep_scale = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
    const bool aligned = a.ldz % 4 == 0 && a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(a.out_f32) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(a.out_bf16) % 8 == 0 &&
                         reinterpret_cast<uintptr_t>(a.out_bf16_lo) % 8 == 0;
    const RowGrid rg = row_grid(a.m, a.width, aligned);
    if (rg.vec == 4)
      AREAD_LAUNCH(bn_act_kernel<4>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale);
    else
      AREAD_LAUNCH(bn_act_kernel<1>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale);
  }
  return AREAD_OK;
}

int aread_bn_act_bwd(const aread_bn_act_bwd_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "bn_act_bwd: null args");
  const aread_bn_act_bwd_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.width > 0 && a.width <= 65536, "bn_act_bwd: width %d unsupported",
                a.width);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE((a.z || a.z_bf16) && a.d_out && a.mean && a.rstd && a.scale && a.shift, "bn_act_bwd: null pointer");
  // row starts of z: 16-byte aligned fp32 rows or 8-byte aligned bf16 rows allow the 4-wide loads
  const bool z_aligned = a.ldz % 4 == 0 && (a.z_bf16 ? reinterpret_cast<uintptr_t>(a.z_bf16) % 8 == 0
                                                     : reinterpret_cast<uintptr_t>(a.z) % 16 == 0);
  AREAD_REQUIRE(a.workspace_bytes >= aread_bn_workspace_bytes(a.width), "bn_act_bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partial = static_cast<float*>(a.workspace);
  float* coef = partial + static_cast<size_t>(kStatCtas) * 2 * a.width;
  const int n_partial = stat_ctas(a.m);
  const bool drop = a.dropout_p > 0.f;
  const uint32_t threshold = drop ? dropout_threshold(a.dropout_p) : 0u;
  const float keep_scale = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  const Geometry g = geometry(a.width, z_aligned && a.ldd % 4 == 0 && reinterpret_cast<uintptr_t>(a.d_out) % 16 == 0);
  const size_t smem = sizeof(float) * 2 * kThreads * g.vec;
  if (g.vec == 4)
    AREAD_LAUNCH(bn_bwd_stats_kernel<4>, n_partial, kThreads, smem, stream, a, g.tw, threshold, keep_scale, partial);
  else
    AREAD_LAUNCH(bn_bwd_stats_kernel<1>, n_partial, kThreads, smem, stream, a, g.tw, threshold, keep_scale, partial);
  AREAD_LAUNCH(bn_bwd_finalize_kernel, ceil_div(a.width, 32), kThreads, 0, stream, a, partial, n_partial, coef);
  if (a.dz_f32 || a.dz_bf16) {
    const bool aligned = z_aligned && a.ldo % 4 == 0 && a.ldd % 4 == 0 &&
                         reinterpret_cast<uintptr_t>(a.d_out) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(a.dz_f32) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(a.dz_bf16) % 8 == 0 &&
                         reinterpret_cast<uintptr_t>(a.dz_bf16_lo) % 8 == 0;
    const RowGrid rg = row_grid(a.m, a.width, aligned);
    if (rg.vec == 4)
      AREAD_LAUNCH(bn_bwd_apply_kernel<4>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale, coef);
    else
      AREAD_LAUNCH(bn_bwd_apply_kernel<1>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale, coef);
  }
  return AREAD_OK;
}

// Activation pass alone, with the statistics (scale / shift) already finalised by the caller.
int aread_bn_act_apply(const aread_bn_act_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "bn_act_apply: null args");
  const aread_bn_act_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.width > 0, "bn_act_apply: bad shape");
  AREAD_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, "bn_act_apply: dropout %f not in [0, 1)", a.dropout_p);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.z && a.scale && a.shift && (a.out_f32 || a.out_bf16), "bn_act_apply: null pointer");
  const bool drop = a.training && a.dropout_p > 0.f;
  const uint32_t threshold = drop ? dropout_threshold(a.dropout_p) : 0u;
  const float keep_scale = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  const bool aligned = a.ldz % 4 == 0 && a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(a.out_f32) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(a.out_bf16) % 8 == 0 &&
                       reinterpret_cast<uintptr_t>(a.out_bf16_lo) % 8 == 0;
  const RowGrid rg = row_grid(a.m, a.width, aligned);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rg.vec == 4)
    AREAD_LAUNCH(bn_act_kernel<4>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale);
  else
    AREAD_LAUNCH(bn_act_kernel<1>, rg.grid, kThreads, 0, stream, a, rg.tw, threshold, keep_scale);
  return AREAD_OK;
}

// Reduction half of the backward alone: d_gamma / d_beta / d_bias and coef = [sum(dy) / m | sum(dy * xhat) / m].
int aread_bn_bwd_coef(const aread_bn_act_bwd_args* args, float* coef, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr && coef != nullptr, "bn_bwd_coef: null args");
  const aread_bn_act_bwd_args& a = *args;
  AREAD_REQUIRE(a.m > 0 && a.width > 0 && a.width <= 65536, "bn_bwd_coef: bad shape");
  AREAD_REQUIRE(a.z && a.d_out && a.mean && a.rstd && a.scale && a.shift, "bn_bwd_coef: null pointer");
  AREAD_REQUIRE(a.workspace_bytes >= aread_bn_workspace_bytes(a.width), "bn_bwd_coef: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partial = static_cast<float*>(a.workspace);
  const int n_partial = stat_ctas(a.m);
  const bool drop = a.dropout_p > 0.f;
  const uint32_t threshold = drop ? dropout_threshold(a.dropout_p) : 0u;
  const float keep_scale = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  const Geometry g = geometry(a.width, a.ldz % 4 == 0 && a.ldd % 4 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0 &&
                                           reinterpret_cast<uintptr_t>(a.d_out) % 16 == 0);
  const size_t smem = sizeof(float) * 2 * kThreads * g.vec;
  if (g.vec == 4)
    AREAD_LAUNCH(bn_bwd_stats_kernel<4>, n_partial, kThreads, smem, stream, a, g.tw, threshold, keep_scale, partial);
  else
    AREAD_LAUNCH(bn_bwd_stats_kernel<1>, n_partial, kThreads, smem, stream, a, g.tw, threshold, keep_scale, partial);
  AREAD_LAUNCH(bn_bwd_finalize_kernel, ceil_div(a.width, 32), kThreads, 0, stream, a, partial, n_partial, coef);
  return AREAD_OK;
}

int aread_mmoe_mix(const aread_mmoe_mix_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "mmoe_mix: null args");
  const aread_mmoe_mix_args& a = *args;
  AREAD_REQUIRE(a.n_expert > 0 && a.n_expert <= 16 && a.n_gate > 0 && a.n_gate <= 8 && a.n_expert * a.n_gate <= 64,
                "mmoe_mix: %d experts x %d gates unsupported", a.n_expert, a.n_gate);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE((a.z || a.z_bf16) && a.scale && a.shift && a.gate, "mmoe_mix: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool drop = a.dropout_p > 0.f;
  const uint32_t threshold = drop ? dropout_threshold(a.dropout_p) : 0u;
  const float keep_scale = drop ? 1.f / (1.f - a.dropout_p) : 1.f;
  // the shipped configuration (4 experts, up to 3 level-0 towers) gets fully unrolled kernels
#define AREAD_MMOE(KERNEL, GRID)                                                                                    \
  do {                                                                                                              \
    if (a.n_expert == 4 && a.n_gate == 3) AREAD_LAUNCH((KERNEL<4, 3>), GRID, kThreads, 0, stream, a, threshold, keep_scale); \
    else if (a.n_expert == 4 && a.n_gate == 2) AREAD_LAUNCH((KERNEL<4, 2>), GRID, kThreads, 0, stream, a, threshold, keep_scale); \
    else if (a.n_expert == 4 && a.n_gate == 1) AREAD_LAUNCH((KERNEL<4, 1>), GRID, kThreads, 0, stream, a, threshold, keep_scale); \
    else AREAD_LAUNCH((KERNEL<0, 0>), GRID, kThreads, 0, stream, a, threshold, keep_scale);                          \
  } while (0)
  if (a.d_out == nullptr) {
    AREAD_REQUIRE(a.out != nullptr, "mmoe_mix: null out");
    const bool vec4 = a.width % 4 == 0 && a.ldz % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(a.scale) % 16 == 0 && reinterpret_cast<uintptr_t>(a.shift) % 16 == 0 &&
                      (a.z_bf16 ? reinterpret_cast<uintptr_t>(a.z_bf16) % 8 == 0 : reinterpret_cast<uintptr_t>(a.z) % 16 == 0);
    if (vec4) {
      AREAD_MMOE(mmoe_mix_fwd_vec4_kernel, elementwise_grid(a.m * (a.width / 4)));
    } else {
      AREAD_MMOE(mmoe_mix_fwd_kernel, elementwise_grid(a.m * a.width));
    }
  } else {
    AREAD_REQUIRE(a.d_h && a.d_gate, "mmoe_mix: null gradient output");
    const bool vec4 = (a.width == 16 || a.width == 32 || a.width == 64 || a.width == 128) && a.ldz % 4 == 0 &&
                      reinterpret_cast<uintptr_t>(a.d_out) % 16 == 0 && reinterpret_cast<uintptr_t>(a.d_h) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(a.scale) % 16 == 0 && reinterpret_cast<uintptr_t>(a.shift) % 16 == 0 &&
                      (a.z_bf16 ? reinterpret_cast<uintptr_t>(a.z_bf16) % 8 == 0 : reinterpret_cast<uintptr_t>(a.z) % 16 == 0);
    if (vec4) {
      const int64_t groups = (a.m + (128 / a.width) - 1) / (128 / a.width);      // samples per warp = 32 / (width / 4)
      AREAD_MMOE(mmoe_mix_bwd_vec4_kernel, elementwise_grid(groups * 32));
    } else {
      AREAD_MMOE(mmoe_mix_bwd_kernel, elementwise_grid(a.m * 32));
    }
  }
#undef AREAD_MMOE
  return AREAD_OK;
}

int aread_dropout_mask(uint64_t seed, uint32_t salt, int64_t n, float p, uint8_t* out, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  if (n == 0) return AREAD_OK;
  AREAD_REQUIRE(out != nullptr, "dropout_mask: null out");
  AREAD_LAUNCH(dropout_mask_kernel, elementwise_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream_), seed, salt, n,
               dropout_threshold(p), out);
  return AREAD_OK;
}

}  // extern "C"
