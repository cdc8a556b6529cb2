// Multi-field embedding lookup (gather) and its gradient (sort-by-row segmented scatter-add).
//
// Reference semantics: model/layer.py:160-183 (forward) and the autograd of layer.py:166
// (aten::embedding_dense_backward).  Both kernels are HBM-bound byte movers: one table row is
// D fp32 = 128 B at the default D = 32, moved by D/4 lanes as one float4 each, so a warp moves
// four rows per instruction and every global access is a full 128-byte line.

#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kTile = AREAD_SCATTER_TILE;

// ---------------------------------------------------------------------------------------------
// plan staging: the per-column / per-field tables are tiny, every CTA copies them to smem once
// ---------------------------------------------------------------------------------------------
struct PlanView {
  const int* col_offset;
  const int* field_src;
  const int* field_nsrc;
  const float* field_div;
};

__device__ __forceinline__ PlanView stage_plan(const aread_embed_plan& p, int* smem) {
  int* s_off = smem;
  int* s_src = s_off + p.n_cols;
  int* s_nsrc = s_src + p.n_fields * p.max_src;
  float* s_div = reinterpret_cast<float*>(s_nsrc + p.n_fields);
  for (int i = threadIdx.x; i < p.n_cols; i += blockDim.x) s_off[i] = p.col_offset[i];
  for (int i = threadIdx.x; i < p.n_fields * p.max_src; i += blockDim.x) s_src[i] = p.field_src[i];
  for (int i = threadIdx.x; i < p.n_fields; i += blockDim.x) {
    s_nsrc[i] = p.field_nsrc[i];
    s_div[i] = p.field_div[i];
  }
  __syncthreads();
  return PlanView{s_off, s_src, s_nsrc, s_div};
}

inline size_t plan_smem_bytes(const aread_embed_plan& p) {
  return sizeof(int) * (static_cast<size_t>(p.n_cols) + static_cast<size_t>(p.n_fields) * p.max_src +
                        2 * static_cast<size_t>(p.n_fields));
}

// idx = x + offset in int32 with wrap-around, exactly like the int32 tensor add of layer.py:165
__device__ __forceinline__ int row_of(int id, int off) {
  return static_cast<int>(static_cast<unsigned>(id) + static_cast<unsigned>(off));
}

__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<unsigned*>(&lo);
  r.y = *reinterpret_cast<unsigned*>(&hi);
  return r;
}

// ---------------------------------------------------------------------------------------------
// gather: LPR lanes (one float4 each) own one sample and walk its output fields, UNROLL fields at
// a time so that every lane group keeps UNROLL independent 128-byte row reads in flight.  All lane
// groups of a warp are in the same field at the same time, so pooled fields do not diverge.
// ---------------------------------------------------------------------------------------------
template <int LPR, int UNROLL>
__global__ void __launch_bounds__(kThreads) gather_kernel(const aread_gather_args a) {
  extern __shared__ int smem[];
  const aread_embed_plan& p = a.plan;
  const PlanView pv = stage_plan(p, smem);

  const int D = p.embed_dim, F = p.n_fields, C = p.n_cols, MS = p.max_src;
  const int lane = threadIdx.x % LPR;
  const bool lane_on = lane * 4 < D;
  const int groups_per_cta = blockDim.x / LPR;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * groups_per_cta;
  const int64_t n_rows = p.n_rows;
  const float* __restrict__ table = a.table + lane * 4;
  int* __restrict__ status = a.status;
  const int shard_shift = a.shard_shift;
  const int shard_mask = (1 << shard_shift) - 1;

  auto fetch = [&](const int* __restrict__ xr, int c, float4& v) {
    const int row = row_of(__ldg(xr + c), pv.col_offset[c]);
    if (row < 0 || row >= n_rows) {
      if (lane == 0 && atomicExch(status, 1) == 0) status[1] = row;
    } else if (lane_on) {
      if (shard_shift == 0) {
        v = ldg4(table + static_cast<int64_t>(row) * D);
      } else {  // the owner's shard, local or peer-mapped over NVLink
        const float* shard = a.shards[row & shard_mask];
        v = *reinterpret_cast<const float4*>(shard + static_cast<int64_t>(row >> shard_shift) * D + lane * 4);
      }
    }
  };

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * groups_per_cta + threadIdx.x / LPR; b < a.batch;
       b += n_groups) {
    const int* __restrict__ xr = a.x + b * C;
    float* __restrict__ orow = a.out + b * F * D + lane * 4;
    uint16_t* __restrict__ hrow = a.out_bf16 ? a.out_bf16 + b * F * D + lane * 4 : nullptr;
    uint16_t* __restrict__ lrow = a.out_bf16_lo ? a.out_bf16_lo + b * F * D + lane * 4 : nullptr;
    for (int f0 = 0; f0 < F; f0 += UNROLL) {
      float4 acc[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f0 + u < F) fetch(xr, pv.field_src[(f0 + u) * MS], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int f = f0 + u;
        if (f < F && pv.field_nsrc[f] > 1) {  // pooled multi-hot field: in-order sum over the positions
          const int nsrc = pv.field_nsrc[f];
          for (int k = 1; k < nsrc; k += 4) {
            float4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k + j < nsrc) fetch(xr, pv.field_src[f * MS + k + j], r[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k + j < nsrc) add4(acc[u], r[j]);
          }
          const float dv = pv.field_div[f];
          if (dv != 1.f) {
            acc[u].x = acc[u].x / dv; acc[u].y = acc[u].y / dv;
            acc[u].z = acc[u].z / dv; acc[u].w = acc[u].w / dv;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int f = f0 + u;
        if (f < F && lane_on) {
          *reinterpret_cast<float4*>(orow + f * D) = acc[u];
          if (hrow != nullptr) {
            const uint2 hi = pack_bf16x4(acc[u]);
            *reinterpret_cast<uint2*>(hrow + f * D) = hi;
            if (lrow != nullptr) {  // split residual: bf16(x - bf16(x))
              const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.x));
              const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.y));
              *reinterpret_cast<uint2*>(lrow + f * D) =
                  pack_bf16x4(make_float4(acc[u].x - h0.x, acc[u].y - h0.y, acc[u].z - h1.x, acc[u].w - h1.y));
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter step 1: (row key, flattened position) pairs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) scatter_keys_kernel(const aread_embed_plan p, const int* __restrict__ x,
                                                                int64_t n, int col_shift,
                                                                unsigned* __restrict__ keys, int* __restrict__ pos) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int b = static_cast<int>(i / p.n_cols);
    const int c = static_cast<int>(i - static_cast<int64_t>(b) * p.n_cols);
    const int row = row_of(x[i], __ldg(p.col_offset + c));
    // out-of-range ids (already reported by the forward) sort behind every real row and are skipped
    keys[i] = (row < 0 || row >= p.n_rows) ? static_cast<unsigned>(p.n_rows) : static_cast<unsigned>(row);
    pos[i] = (b << col_shift) | c;  // (sample, column) packed so that the reduce needs no division
  }
}

// sorted payloads back to flattened positions b * n_cols + c (only for callers that ask for them)
__global__ void __launch_bounds__(kThreads) scatter_unpack_kernel(const int* __restrict__ packed, int64_t n,
                                                                  int col_shift, int n_cols, int* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int q = packed[i];
    out[i] = (q >> col_shift) * n_cols + (q & ((1 << col_shift) - 1));
  }
}

// ---------------------------------------------------------------------------------------------
// scatter step 3: one lane group per tile of kTile sorted lookups, in-order run sums.  The tile's
// keys / positions are loaded once (kTile / LPR per lane) and handed round by shuffles; gradient
// rows are fetched eight at a time so the reads overlap, then consumed strictly in order.
// ---------------------------------------------------------------------------------------------
// row of the (possibly owner-major) gradient buffer that receives table row `key`
__device__ __forceinline__ int64_t dest_row(unsigned key, int shard_shift, int64_t shard_rows) {
  return shard_shift == 0 ? static_cast<int64_t>(key)
                          : static_cast<int64_t>(key & ((1u << shard_shift) - 1u)) * shard_rows + (key >> shard_shift);
}

// where the finished sum of table row `key` goes: the (possibly owner-major) local buffer, or -- sparse exchange --
// the owning rank's receive buffer for this sender (a peer-mapped address: the store travels over NVLink)
__device__ __forceinline__ float* dest_ptr(unsigned key, int D, float* d_table, float* const* __restrict__ peer_grads,
                                           int shard_shift, int64_t shard_rows) {
  if (peer_grads != nullptr)
    return peer_grads[key & ((1u << shard_shift) - 1u)] + static_cast<int64_t>(key >> shard_shift) * D;
  return d_table + dest_row(key, shard_shift, shard_rows) * D;
}

template <int LPR>
__global__ void __launch_bounds__(kThreads) scatter_tile_kernel(
    const aread_embed_plan p, const unsigned* __restrict__ keys, const int* __restrict__ pos, int64_t n,
    int64_t n_tiles, int col_shift, const float* __restrict__ d_out, float* __restrict__ d_table,
    float* __restrict__ carry_in, float* __restrict__ carry_out, int shard_shift, int64_t shard_rows,
    float* const* __restrict__ peer_grads) {
  extern __shared__ int smem[];
  int* s_field = smem;                                        // column -> output field
  float* s_div = reinterpret_cast<float*>(smem + p.n_cols);   // column -> fl32(1 / pooling divisor)
  for (int i = threadIdx.x; i < p.n_fields * p.max_src; i += blockDim.x) {
    const int f = i / p.max_src, k = i - f * p.max_src;
    if (k < p.field_nsrc[f]) {
      const int c = p.field_src[i];
      s_field[c] = f;
      s_div[c] = 1.f / p.field_div[f];
    }
  }
  __syncthreads();

  constexpr int PER = kTile / LPR;  // tile entries held per lane
  constexpr int BATCH = 8;
  const int D = p.embed_dim, F = p.n_fields;
  const int col_mask = (1 << col_shift) - 1;
  const int lane = threadIdx.x % LPR;
  const bool lane_on = lane * 4 < D;
  const unsigned gmask =
      LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (((threadIdx.x % 32) / LPR) * LPR));
  const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x / LPR) + threadIdx.x / LPR;
  if (t >= n_tiles) return;
  const int64_t start = t * kTile;
  const int cnt = static_cast<int>(min(n - start, static_cast<int64_t>(kTile)));
  const unsigned n_rows = static_cast<unsigned>(p.n_rows);

  unsigned k_reg[PER];
  int q_reg[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int e = j * LPR + lane;
    k_reg[j] = e < cnt ? __ldg(keys + start + e) : 0xffffffffu;
    q_reg[j] = e < cnt ? __ldg(pos + start + e) : 0;
  }
  const float* __restrict__ g_base = d_out + lane * 4;

  auto flush = [&](unsigned key, const float4& acc, bool leading, bool trailing) {
    if (key >= n_rows) return;
    if (leading) {
      if (lane_on) *reinterpret_cast<float4*>(carry_in + t * D + lane * 4) = acc;
    } else if (trailing) {
      if (lane_on) *reinterpret_cast<float4*>(carry_out + t * D + lane * 4) = acc;
    } else if (lane_on) {
      *reinterpret_cast<float4*>(dest_ptr(key, D, d_table, peer_grads, shard_shift, shard_rows) + lane * 4) = acc;
    }
  };

  unsigned cur = 0;
  bool leading = false;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int e0 = 0; e0 < kTile; e0 += BATCH) {
    float4 g[BATCH];
    unsigned kk[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int e = e0 + j;
      kk[j] = __shfl_sync(gmask, k_reg[e / LPR], e % LPR, LPR);
      const int q = __shfl_sync(gmask, q_reg[e / LPR], e % LPR, LPR);
      g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < cnt) {
        const int b = q >> col_shift, c = q & col_mask;
        if (lane_on) g[j] = ldg4(g_base + (static_cast<int64_t>(b) * F + s_field[c]) * D);
        const float sc = s_div[c];
        if (sc != 1.f) { g[j].x *= sc; g[j].y *= sc; g[j].z *= sc; g[j].w *= sc; }
      }
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int e = e0 + j;
      if (e >= cnt) continue;
      if (e == 0) {
        cur = kk[j];
        acc = g[j];
        leading = start > 0 && __ldg(keys + start - 1) == cur;
      } else if (kk[j] == cur) {
        add4(acc, g[j]);
      } else {
        flush(cur, acc, leading, false);
        leading = false;
        cur = kk[j];
        acc = g[j];
      }
    }
  }
  flush(cur, acc, leading, start + cnt < n && __ldg(keys + start + cnt) == cur);
}

// ---------------------------------------------------------------------------------------------
// scatter step 4: rows spanning several tiles.  Blocks of 32^l sorted entries form an aligned
// 32-ary tree over the sorted list.  A level-l block leaves at most two open partial sums behind:
// `in` (its leading run, when that run continues from the previous block) and `out` (its trailing
// run, when that run continues into the next block and is not the leading run).  One lane group
// per parent block walks its 32 children's open partials in order, adds runs of equal rows left to
// right, writes rows that are now complete and passes the still-open ones up.  Serial chains are
// at most 64 adds per level whatever the skew (a row hit by every sample needs log32(n) levels).
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(kThreads) scatter_level_kernel(
    int D, unsigned n_rows, const unsigned* __restrict__ keys, int64_t n, int64_t child_size, int64_t n_children,
    int64_t n_blocks, const float* __restrict__ in_c, const float* __restrict__ out_c, float* __restrict__ in_p,
    float* __restrict__ out_p, float* __restrict__ d_table, int shard_shift, int64_t shard_rows,
    float* const* __restrict__ peer_grads) {
  constexpr int PER = 32 / LPR;  // children held per lane
  constexpr int BATCH = 4;       // children whose partials are fetched together
  const int lane = threadIdx.x % LPR;
  const bool lane_on = lane * 4 < D;
  const unsigned gmask =
      LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (((threadIdx.x % 32) / LPR) * LPR));
  const int64_t J = static_cast<int64_t>(blockIdx.x) * (blockDim.x / LPR) + threadIdx.x / LPR;
  if (J >= n_blocks) return;
  const int64_t c0 = J * 32;

  unsigned kf[PER], kl[PER], fl[PER];  // first key, last key, flags (1: has in, 2: has out)
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int64_t c = c0 + j * LPR + lane;
    const int64_t s = c * child_size;
    kf[j] = kl[j] = 0xffffffffu;
    fl[j] = 0;
    if (c < n_children) {
      const int64_t e = min(n, s + child_size);
      kf[j] = __ldg(keys + s);
      kl[j] = __ldg(keys + e - 1);
      const bool lead = s > 0 && __ldg(keys + s - 1) == kf[j];
      const bool trail = e < n && __ldg(keys + e) == kl[j];
      fl[j] = (lead ? 1u : 0u) | ((trail && !(lead && kf[j] == kl[j])) ? 2u : 0u);
    }
  }
  const int64_t end_J = min(n, (c0 + 32) * child_size);
  const bool trailing_J = end_J < n && __ldg(keys + end_J) == __ldg(keys + end_J - 1);

  unsigned cur = 0;
  bool started = false, leading = false;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&](bool trailing) {
    if (cur >= n_rows) return;
    float* dst = leading ? in_p + J * D
                         : (trailing ? out_p + J * D : dest_ptr(cur, D, d_table, peer_grads, shard_shift, shard_rows));
    if (lane_on) *reinterpret_cast<float4*>(dst + lane * 4) = acc;
  };
  auto element = [&](unsigned key, const float4& v, bool first_of_block) {
    if (!started) {
      started = true;
      cur = key;
      acc = v;
      leading = first_of_block;
    } else if (key == cur) {
      add4(acc, v);
    } else {
      flush(false);
      leading = false;
      cur = key;
      acc = v;
    }
  };

#pragma unroll
  for (int i0 = 0; i0 < 32; i0 += BATCH) {
    unsigned a_kf[BATCH], a_kl[BATCH], a_fl[BATCH];
    float4 v_in[BATCH], v_out[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = i0 + j;
      a_kf[j] = __shfl_sync(gmask, kf[i / LPR], i % LPR, LPR);
      a_kl[j] = __shfl_sync(gmask, kl[i / LPR], i % LPR, LPR);
      a_fl[j] = __shfl_sync(gmask, fl[i / LPR], i % LPR, LPR);
      v_in[j] = v_out[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((a_fl[j] & 1u) && lane_on) v_in[j] = ldg4(in_c + (c0 + i) * D + lane * 4);
      if ((a_fl[j] & 2u) && lane_on) v_out[j] = ldg4(out_c + (c0 + i) * D + lane * 4);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      if (a_fl[j] & 1u) element(a_kf[j], v_in[j], i0 + j == 0);
      if (a_fl[j] & 2u) element(a_kl[j], v_out[j], false);
    }
  }
  if (started) flush(trailing_J);
}

int lanes_per_row(int D) {
  int lpr = 1;
  while (lpr * 4 < D) lpr <<= 1;
  return lpr;
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (row, position) pairs by row: 1 or 3 passes (always odd, so the result lands in the
// "out" buffers) over digits of up to 11 bits.  The unit of work is one WARP (a 32-thread CTA) that owns a
// contiguous chunk of the input and walks it 32 elements at a time, so the order inside a chunk is the input
// order by construction and no block-wide synchronisation is needed:
//   radix_hist_kernel     per-unit digit counts                          -> counts[digit][unit]
//   radix_scan_units / _digits  exclusive prefix over (digit-major, unit-minor) -> where each unit's run of a digit starts
//   radix_scatter_kernel  per element: match_any over the warp gives its rank among equal digits of the round
// Equal rows therefore keep their input order (sample-major), which fixes the accumulation order of the scatter.
constexpr int kSortMaxDigitBits = 11;
constexpr int kSortMaxCounts = 2 * 1024 * 1024;  // digit values x units (8 MB of counters)
constexpr int kSortMaxUnits = kNumSMs * 16;

struct SortPlan {
  int passes, digit_bits, radix, units;
  int64_t per_unit;
};

SortPlan sort_plan(int64_t n, int bits) {
  SortPlan p;
  p.passes = bits <= kSortMaxDigitBits ? 1 : 3;
  p.digit_bits = p.passes == 1 ? bits : (bits + 2) / 3;
  p.radix = 1 << p.digit_bits;
  int64_t units = (n + 255) / 256;                       // at least 8 rounds per unit
  if (units > kSortMaxUnits) units = kSortMaxUnits;
  if (units > kSortMaxCounts / p.radix) units = kSortMaxCounts / p.radix;
  if (units < 1) units = 1;
  p.units = static_cast<int>(units);
  p.per_unit = ((n + units - 1) / units + 31) / 32 * 32;
  return p;
}

__global__ void __launch_bounds__(32) radix_hist_kernel(const unsigned* __restrict__ keys, int64_t n, int64_t per_unit,
                                                        int shift, int radix, int* __restrict__ counts) {
  extern __shared__ int s_hist[];
  for (int i = threadIdx.x; i < radix; i += 32) s_hist[i] = 0;
  __syncwarp();
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * per_unit;
  const int64_t end = begin + per_unit < n ? begin + per_unit : n;
  for (int64_t i = begin + threadIdx.x; i < end; i += 32) atomicAdd(&s_hist[(__ldg(keys + i) >> shift) & (radix - 1)], 1);
  __syncwarp();
  for (int i = threadIdx.x; i < radix; i += 32) counts[static_cast<int64_t>(i) * gridDim.x + blockIdx.x] = s_hist[i];
}

// counts[digit][unit] -> exclusive prefix over the units of that digit (one warp per digit, coalesced), digit totals
__global__ void __launch_bounds__(256) radix_scan_units_kernel(int* __restrict__ counts, int radix, int units,
                                                                int* __restrict__ totals) {
  const int digit = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (digit >= radix) return;
  int* row = counts + static_cast<int64_t>(digit) * units;
  int carry = 0;
  for (int u0 = 0; u0 < units; u0 += 32) {
    const int u = u0 + lane;
    const int c = u < units ? row[u] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (u < units) row[u] = carry + incl - c;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) totals[digit] = carry;
}

// totals[digit] -> exclusive prefix over the digits (radix <= 2048: one CTA, two values per thread)
__global__ void __launch_bounds__(1024) radix_scan_digits_kernel(int* __restrict__ totals, int radix) {
  __shared__ int s_part[1024];
  const int i0 = threadIdx.x * 2;
  const int a = i0 < radix ? totals[i0] : 0, b = i0 + 1 < radix ? totals[i0 + 1] : 0;
  s_part[threadIdx.x] = a + b;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  const int excl = s_part[threadIdx.x] - (a + b);
  if (i0 < radix) totals[i0] = excl;
  if (i0 + 1 < radix) totals[i0 + 1] = excl + a;
}

__global__ void __launch_bounds__(32) radix_scatter_kernel(const unsigned* __restrict__ keys_in,
                                                           const int* __restrict__ vals_in, int64_t n, int64_t per_unit,
                                                           int shift, int radix, const int* __restrict__ offsets,
                                                           const int* __restrict__ digit_base,
                                                           unsigned* __restrict__ keys_out, int* __restrict__ vals_out) {
  extern __shared__ int s_base[];
  for (int i = threadIdx.x; i < radix; i += 32)
    s_base[i] = digit_base[i] + offsets[static_cast<int64_t>(i) * gridDim.x + blockIdx.x];
  __syncwarp();
  const int lane = threadIdx.x;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * per_unit;
  const int64_t end = begin + per_unit < n ? begin + per_unit : n;
  for (int64_t i0 = begin; i0 < end; i0 += 32) {
    const int64_t i = i0 + lane;
    const bool valid = i < end;
    unsigned key = 0;
    int val = 0;
    if (valid) {
      key = __ldg(keys_in + i);
      val = __ldg(vals_in + i);
    }
    const int digit = valid ? static_cast<int>((key >> shift) & (radix - 1)) : radix;   // padding lanes: own class
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    int start = 0;
    if (valid && lane == leader) {
      start = s_base[digit];
      s_base[digit] = start + __popc(peers);
    }
    start = __shfl_sync(0xffffffffu, start, leader);
    if (valid) {
      const int dst = start + __popc(peers & lt_mask);
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
    __syncwarp();
  }
}

int radix_sort_pairs(unsigned* keys_in, unsigned* keys_out, int* vals_in, int* vals_out, int64_t n, int bits, int* counts,
                     cudaStream_t stream) {
  const SortPlan p = sort_plan(n, bits);
  const size_t smem = sizeof(int) * p.radix;
  int* totals = counts + static_cast<int64_t>(p.radix) * p.units;     // [radix], after the counts matrix
  unsigned *ki = keys_in, *ko = keys_out;
  int *vi = vals_in, *vo = vals_out;
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.digit_bits;
    AREAD_LAUNCH(radix_hist_kernel, p.units, 32, smem, stream, ki, n, p.per_unit, shift, p.radix, counts);
    AREAD_LAUNCH(radix_scan_units_kernel, ceil_div(p.radix, 8), 256, 0, stream, counts, p.radix, p.units, totals);
    AREAD_LAUNCH(radix_scan_digits_kernel, 1, 1024, 0, stream, totals, p.radix);
    AREAD_LAUNCH(radix_scatter_kernel, p.units, 32, smem, stream, ki, vi, n, p.per_unit, shift, p.radix, counts, totals, ko,
                 vo);
    unsigned* tk = ki; ki = ko; ko = tk;
    int* tv = vi; vi = vo; vo = tv;
  }
  return AREAD_OK;   // passes is odd: the sorted pairs are in keys_out / vals_out
}

int key_bits(int64_t n_rows) {  // bits needed for keys 0 .. n_rows (inclusive: the skip sentinel)
  int bits = 1;
  while ((int64_t{1} << bits) <= n_rows) ++bits;
  return bits;
}

struct ScatterWorkspace {
  unsigned* keys_in;
  unsigned* keys_out;
  int* pos_in;
  int* pos_out;
  float* in_a;   // open partials of the odd levels (sized for the tiles)
  float* out_a;
  float* in_b;   // open partials of the even levels (sized for tiles / 32)
  float* out_b;
  int* sort_counts;   // radix_sort_pairs: digit values x units
  size_t total;
};

int carve_scatter_workspace(void* base, int64_t n, int D, ScatterWorkspace* w) {
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t n_l2 = (n_tiles + 31) / 32;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  w->keys_in = reinterpret_cast<unsigned*>(take(nn * 4));
  w->keys_out = reinterpret_cast<unsigned*>(take(nn * 4));
  w->pos_in = reinterpret_cast<int*>(take(nn * 4));
  w->pos_out = reinterpret_cast<int*>(take(nn * 4));
  w->in_a = reinterpret_cast<float*>(take(static_cast<size_t>(n_tiles + 1) * D * 4));
  w->out_a = reinterpret_cast<float*>(take(static_cast<size_t>(n_tiles + 1) * D * 4));
  w->in_b = reinterpret_cast<float*>(take(static_cast<size_t>(n_l2 + 1) * D * 4));
  w->out_b = reinterpret_cast<float*>(take(static_cast<size_t>(n_l2 + 1) * D * 4));
  w->sort_counts = reinterpret_cast<int*>(take((static_cast<size_t>(kSortMaxCounts) + (1 << kSortMaxDigitBits)) * 4));
  w->total = off;
  return AREAD_OK;
}

int check_plan(const aread_embed_plan& p) {
  AREAD_REQUIRE(p.n_cols > 0 && p.n_fields > 0 && p.max_src > 0, "embed plan: empty layout");
  AREAD_REQUIRE(p.embed_dim > 0 && p.embed_dim % 4 == 0 && p.embed_dim <= 128,
                "embed plan: embed_dim %d must be a multiple of 4 and <= 128", p.embed_dim);
  AREAD_REQUIRE(p.n_rows > 0 && p.n_rows < (int64_t{1} << 31), "embed plan: n_rows %lld out of range",
                static_cast<long long>(p.n_rows));
  AREAD_REQUIRE(p.col_offset && p.field_src && p.field_nsrc && p.field_div, "embed plan: null table");
  return AREAD_OK;
}

template <int LPR>
int launch_gather(const aread_gather_args& a, cudaStream_t stream) {
  constexpr int kUnroll = 4;
  int64_t grid = (a.batch * LPR + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;  // 8 CTAs of 256 threads fill an SM
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  AREAD_LAUNCH((gather_kernel<LPR, kUnroll>), static_cast<unsigned>(grid), kThreads, plan_smem_bytes(a.plan), stream,
               a);
  return AREAD_OK;
}

template <int LPR>
int launch_scatter(const aread_scatter_args& a, const ScatterWorkspace& w, int64_t n, int col_shift,
                   cudaStream_t stream) {
  const int D = a.plan.embed_dim;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int groups_per_cta = kThreads / LPR;
  const size_t smem = sizeof(int) * 2 * static_cast<size_t>(a.plan.n_cols);
  AREAD_LAUNCH((scatter_tile_kernel<LPR>), static_cast<unsigned>((n_tiles + groups_per_cta - 1) / groups_per_cta),
               kThreads, smem, stream, a.plan, w.keys_out, w.pos_out, n, n_tiles, col_shift, a.d_out, a.d_table,
               w.in_a, w.out_a, a.shard_shift, a.shard_rows, a.peer_grads);
  const float *in_c = w.in_a, *out_c = w.out_a;
  float *in_p = w.in_b, *out_p = w.out_b;
  int64_t child_size = kTile, n_children = n_tiles;
  while (n_children > 1) {
    const int64_t n_blocks = (n_children + 31) / 32;
    AREAD_LAUNCH((scatter_level_kernel<LPR>), static_cast<unsigned>((n_blocks + groups_per_cta - 1) / groups_per_cta),
                 kThreads, 0, stream, D, static_cast<unsigned>(a.plan.n_rows), w.keys_out, n, child_size, n_children,
                 n_blocks, in_c, out_c, in_p, out_p, a.d_table, a.shard_shift, a.shard_rows, a.peer_grads);
    const float* t_in = in_c;
    const float* t_out = out_c;
    in_c = in_p;
    out_c = out_p;
    in_p = const_cast<float*>(t_in);
    out_p = const_cast<float*>(t_out);
    child_size *= 32;
    n_children = n_blocks;
  }
  return AREAD_OK;
}

// out[i] = scale * sum_s recv[s][i]; non-zero inputs are cleared for the next step
__global__ void __launch_bounds__(kThreads) shard_grad_sum_kernel(float* __restrict__ recv, int n_senders, int64_t n4,
                                                                  float scale, float* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = zero;
    for (int s = 0; s < n_senders; ++s) {
      float4* src = reinterpret_cast<float4*>(recv) + static_cast<int64_t>(s) * n4 + i;
      const float4 v = *src;
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        add4(acc, v);
        *src = zero;
      }
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_shard_grad_sum(float* recv, int32_t n_senders, int64_t n, float scale, float* out, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(recv && out && n_senders > 0 && n >= 0 && n % 4 == 0, "shard_grad_sum: bad arguments");
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(recv) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "shard_grad_sum: buffers must be 16-byte aligned");
  if (n == 0) return AREAD_OK;
  int64_t grid = (n / 4 + kThreads - 1) / kThreads;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  AREAD_LAUNCH(shard_grad_sum_kernel, static_cast<unsigned>(grid), kThreads, 0, static_cast<cudaStream_t>(stream_), recv,
               n_senders, n / 4, scale, out);
  return AREAD_OK;
}

int aread_gather_fwd(const aread_gather_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "gather: null args");
  const aread_gather_args& a = *args;
  if (int rc = check_plan(a.plan)) return rc;
  AREAD_REQUIRE(a.batch >= 0, "gather: negative batch");
  if (a.batch == 0) return AREAD_OK;
  AREAD_REQUIRE(a.x && a.out && a.status, "gather: null pointer");
  AREAD_REQUIRE(a.shard_shift >= 0 && a.shard_shift <= 6, "gather: shard_shift %d out of range", a.shard_shift);
  AREAD_REQUIRE(a.shard_shift == 0 ? a.table != nullptr : a.shards != nullptr, "gather: null table");
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(a.table) | reinterpret_cast<uintptr_t>(a.out)) % 16 == 0,
                "gather: table/out must be 16-byte aligned");
  AREAD_REQUIRE(a.out_bf16 == nullptr || reinterpret_cast<uintptr_t>(a.out_bf16) % 8 == 0,
                "gather: out_bf16 must be 8-byte aligned");
  AREAD_REQUIRE(a.out_bf16_lo == nullptr || (a.out_bf16 != nullptr && reinterpret_cast<uintptr_t>(a.out_bf16_lo) % 8 == 0),
                "gather: out_bf16_lo needs out_bf16 and 8-byte alignment");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (lanes_per_row(a.plan.embed_dim)) {
    case 1: return launch_gather<1>(a, stream);
    case 2: return launch_gather<2>(a, stream);
    case 4: return launch_gather<4>(a, stream);
    case 8: return launch_gather<8>(a, stream);
    case 16: return launch_gather<16>(a, stream);
    default: return launch_gather<32>(a, stream);
  }
}

size_t aread_scatter_workspace_bytes(int64_t n_lookups, int32_t embed_dim) {
  aread::ScatterWorkspace w;
  if (aread::carve_scatter_workspace(nullptr, n_lookups, embed_dim, &w) != AREAD_OK) return 0;
  return w.total;
}

int aread_scatter_bwd(const aread_scatter_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "scatter: null args");
  const aread_scatter_args& a = *args;
  if (int rc = check_plan(a.plan)) return rc;
  AREAD_REQUIRE(a.batch >= 0, "scatter: negative batch");
  AREAD_REQUIRE(a.d_table != nullptr || a.peer_grads != nullptr, "scatter: null d_table");
  AREAD_REQUIRE(a.peer_grads == nullptr || a.shard_shift > 0, "scatter: peer_grads needs a sharded table");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D = a.plan.embed_dim;
  AREAD_REQUIRE(a.shard_shift >= 0 && a.shard_shift <= 6 && (a.shard_shift == 0 || a.shard_rows > 0),
                "scatter: bad shard layout");
  const size_t grad_rows = a.shard_shift == 0 ? static_cast<size_t>(a.plan.n_rows)
                                              : static_cast<size_t>(a.shard_rows) << a.shard_shift;
  if (a.zero_fill && a.peer_grads == nullptr)
    AREAD_CUDA(cudaMemsetAsync(a.d_table, 0, grad_rows * D * sizeof(float), stream));
  const int64_t n = a.batch * a.plan.n_cols;
  if (n == 0) return AREAD_OK;
  int col_shift = 0;
  while ((1 << col_shift) < a.plan.n_cols) ++col_shift;
  AREAD_REQUIRE((a.batch << col_shift) < (int64_t{1} << 31), "scatter: batch %lld x %d columns exceeds int32 positions",
                static_cast<long long>(a.batch), a.plan.n_cols);
  AREAD_REQUIRE(a.x && a.d_out && a.workspace, "scatter: null pointer");
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(a.d_table) | reinterpret_cast<uintptr_t>(a.d_out) |
                 reinterpret_cast<uintptr_t>(a.workspace)) % 16 == 0,
                "scatter: d_table/d_out/workspace must be 16-byte aligned");
  ScatterWorkspace w;
  if (int rc = carve_scatter_workspace(a.workspace, n, D, &w)) return rc;
  if (w.total > a.workspace_bytes)
    return fail(AREAD_ERR_WORKSPACE, "scatter: workspace %zu < %zu bytes", a.workspace_bytes, w.total);

  {
    int64_t grid = (n + kThreads - 1) / kThreads;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    AREAD_LAUNCH(scatter_keys_kernel, static_cast<unsigned>(grid), kThreads, 0, stream, a.plan, a.x, n, col_shift,
                 w.keys_in, w.pos_in);
  }
  // Stable LSD radix sort by table row; the payload is the flattened (sample, column) position, so equal rows keep
  // the reference's accumulation order.
  if (int rc = radix_sort_pairs(w.keys_in, w.keys_out, w.pos_in, w.pos_out, n, key_bits(a.plan.n_rows), w.sort_counts,
                                stream))
    return rc;
  int rc;
  switch (lanes_per_row(D)) {
    case 1: rc = launch_scatter<1>(a, w, n, col_shift, stream); break;
    case 2: rc = launch_scatter<2>(a, w, n, col_shift, stream); break;
    case 4: rc = launch_scatter<4>(a, w, n, col_shift, stream); break;
    case 8: rc = launch_scatter<8>(a, w, n, col_shift, stream); break;
    case 16: rc = launch_scatter<16>(a, w, n, col_shift, stream); break;
    default: rc = launch_scatter<32>(a, w, n, col_shift, stream); break;
  }
  if (rc) return rc;
  if (a.sorted_rows)
    AREAD_CUDA(cudaMemcpyAsync(a.sorted_rows, w.keys_out, n * 4, cudaMemcpyDeviceToDevice, stream));
  if (a.sorted_pos) {
    int64_t grid = (n + kThreads - 1) / kThreads;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    AREAD_LAUNCH(scatter_unpack_kernel, static_cast<unsigned>(grid), kThreads, 0, stream, w.pos_out, n, col_shift,
                 a.plan.n_cols, a.sorted_pos);
  }
  return AREAD_OK;
}

}  // extern "C"
