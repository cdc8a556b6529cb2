// Multi-field embedding lookup (gather) and its gradient (sort-by-row segmented scatter-add).
//
// Reference semantics: model/layer.py:160-183 (forward) and the autograd of layer.py:166
// (aten::embedding_dense_backward).  Both kernels are HBM-bound byte movers: one table row is
// D fp32 = 128 B at the default D = 32, moved by D/4 lanes as one float4 each, so a warp moves
// four rows per instruction and every global access is a full 128-byte line.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace aread {
namespace {

constexpr int kThreads = 256;
constexpr int kTile = AREAD_SCATTER_TILE;
constexpr int kSpanBlocks = AREAD_SCATTER_SPAN_BLOCKS;  // fixed: part of the documented summation order

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
// gather: one "slot" = one (sample, output field); LPR lanes move one slot's D floats.
// Each lane group keeps UNROLL independent slots in flight to cover the HBM latency.
// ---------------------------------------------------------------------------------------------
template <int LPR, int UNROLL>
__global__ void __launch_bounds__(kThreads) gather_kernel(const aread_gather_args a) {
  extern __shared__ int smem[];
  const aread_embed_plan& p = a.plan;
  const PlanView pv = stage_plan(p, smem);

  const int D = p.embed_dim;
  const int F = p.n_fields;
  const int C = p.n_cols;
  const int lane = threadIdx.x % LPR;
  const bool lane_on = lane * 4 < D;
  const int64_t groups_per_cta = blockDim.x / LPR;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * groups_per_cta;
  const int64_t group = static_cast<int64_t>(blockIdx.x) * groups_per_cta + threadIdx.x / LPR;
  const int64_t n_slots = a.batch * F;
  const int64_t n_rows = p.n_rows;

  for (int64_t s0 = group; s0 < n_slots; s0 += n_groups * UNROLL) {
    float4 acc[UNROLL];
    int64_t b[UNROLL];
    int f[UNROLL];
    int nsrc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t s = s0 + u * n_groups;
      acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      nsrc[u] = 0;
      b[u] = 0;
      f[u] = 0;
      if (s < n_slots) {
        b[u] = s / F;
        f[u] = static_cast<int>(s - b[u] * F);
        nsrc[u] = pv.field_nsrc[f[u]];
        const int c = pv.field_src[f[u] * p.max_src];
        const int row = row_of(__ldg(a.x + b[u] * C + c), pv.col_offset[c]);
        if (row < 0 || row >= n_rows) {
          if (lane == 0 && atomicExch(a.status, 1) == 0) a.status[1] = row;
        } else if (lane_on) {
          acc[u] = ldg4(a.table + static_cast<int64_t>(row) * D + lane * 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (nsrc[u] > 1) {  // pooled multi-hot field: in-order sum over the sequence positions
        for (int k = 1; k < nsrc[u]; ++k) {
          const int c = pv.field_src[f[u] * p.max_src + k];
          const int row = row_of(__ldg(a.x + b[u] * C + c), pv.col_offset[c]);
          if (row < 0 || row >= n_rows) {
            if (lane == 0 && atomicExch(a.status, 1) == 0) a.status[1] = row;
          } else if (lane_on) {
            add4(acc[u], ldg4(a.table + static_cast<int64_t>(row) * D + lane * 4));
          }
        }
        const float dv = pv.field_div[f[u]];
        if (dv != 1.f) {
          acc[u].x = acc[u].x / dv; acc[u].y = acc[u].y / dv;
          acc[u].z = acc[u].z / dv; acc[u].w = acc[u].w / dv;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t s = s0 + u * n_groups;
      if (s < n_slots && lane_on) {
        *reinterpret_cast<float4*>(a.out + s * D + lane * 4) = acc[u];
        if (a.out_bf16 != nullptr)
          *reinterpret_cast<uint2*>(a.out_bf16 + s * D + lane * 4) = pack_bf16x4(acc[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter step 1: (row key, flattened position) pairs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) scatter_keys_kernel(const aread_embed_plan p, const int* __restrict__ x,
                                                                int64_t n, unsigned* __restrict__ keys,
                                                                int* __restrict__ pos) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % p.n_cols);
    const int row = row_of(x[i], __ldg(p.col_offset + c));
    // out-of-range ids (already reported by the forward) sort behind every real row and are skipped
    keys[i] = (row < 0 || row >= p.n_rows) ? static_cast<unsigned>(p.n_rows) : static_cast<unsigned>(row);
    pos[i] = static_cast<int>(i);
  }
}

// ---------------------------------------------------------------------------------------------
// scatter step 3: one lane group per tile of kTile sorted lookups, in-order run sums
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(kThreads) scatter_tile_kernel(
    const aread_embed_plan p, const unsigned* __restrict__ keys, const int* __restrict__ pos, int64_t n,
    int64_t n_tiles, const float* __restrict__ d_out, float* __restrict__ d_table, float* __restrict__ carry_in,
    float* __restrict__ carry_out, int* __restrict__ span_count, int* __restrict__ span_tiles) {
  extern __shared__ int smem[];
  int* s_field = smem;                                        // column -> output field
  float* s_div = reinterpret_cast<float*>(smem + p.n_cols);   // column -> pooling divisor
  for (int i = threadIdx.x; i < p.n_fields * p.max_src; i += blockDim.x) {
    const int f = i / p.max_src, k = i - f * p.max_src;
    if (k < p.field_nsrc[f]) {
      const int c = p.field_src[i];
      s_field[c] = f;
      s_div[c] = p.field_div[f];
    }
  }
  __syncthreads();

  const int D = p.embed_dim, F = p.n_fields, C = p.n_cols;
  const int lane = threadIdx.x % LPR;
  const bool lane_on = lane * 4 < D;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x / LPR) + threadIdx.x / LPR;
  if (t >= n_tiles) return;
  const int64_t start = t * kTile;
  const int64_t end = min(n, start + static_cast<int64_t>(kTile));
  const unsigned n_rows = static_cast<unsigned>(p.n_rows);

  auto grad = [&](int64_t i) -> float4 {
    const int q = __ldg(pos + i);
    const int b = q / C, c = q - b * C;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_on) v = ldg4(d_out + (static_cast<int64_t>(b) * F + s_field[c]) * D + lane * 4);
    const float dv = s_div[c];
    if (dv != 1.f) { v.x = v.x / dv; v.y = v.y / dv; v.z = v.z / dv; v.w = v.w / dv; }
    return v;
  };
  auto flush = [&](unsigned key, const float4& acc, bool leading, bool trailing) {
    if (key >= n_rows) return;
    if (leading) {
      if (lane_on) *reinterpret_cast<float4*>(carry_in + t * D + lane * 4) = acc;
    } else if (trailing) {
      if (lane_on) *reinterpret_cast<float4*>(carry_out + t * D + lane * 4) = acc;
      if (lane == 0) span_tiles[atomicAdd(span_count, 1)] = static_cast<int>(t);
    } else if (lane_on) {
      *reinterpret_cast<float4*>(d_table + static_cast<int64_t>(key) * D + lane * 4) = acc;
    }
  };

  unsigned cur = __ldg(keys + start);
  bool leading = start > 0 && __ldg(keys + start - 1) == cur;
  float4 acc = grad(start);
#pragma unroll 4
  for (int64_t i = start + 1; i < end; ++i) {
    const unsigned k = __ldg(keys + i);
    const float4 g = grad(i);
    if (k == cur) {
      add4(acc, g);
    } else {
      flush(cur, acc, leading, false);
      leading = false;
      cur = k;
      acc = g;
    }
  }
  flush(cur, acc, leading, end < n && __ldg(keys + end) == cur);
}

// ---------------------------------------------------------------------------------------------
// scatter step 4: rows spanning several tiles.  One CTA per such row: its K carry-in partials are
// cut into kSpanBlocks contiguous blocks of ceil(K / kSpanBlocks), one lane group sums each block
// left to right, then the block sums are added left to right onto the first tile's carry-out.
// ---------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(kSpanBlocks* LPR) scatter_span_kernel(
    int D, const unsigned* __restrict__ keys, int64_t n, const float* __restrict__ carry_in,
    const float* __restrict__ carry_out, const int* __restrict__ span_count, const int* __restrict__ span_tiles,
    float* __restrict__ d_table) {
  extern __shared__ float s_part[];  // [kSpanBlocks][D]
  __shared__ int64_t s_last_tile;
  const int lane = threadIdx.x % LPR;
  const int blk = threadIdx.x / LPR;
  const bool lane_on = lane * 4 < D;
  const int n_span = *span_count;
  for (int w = blockIdx.x; w < n_span; w += gridDim.x) {
    const int64_t t = span_tiles[w];
    const int64_t t_end = min(n, (t + 1) * kTile);
    const unsigned key = keys[t_end - 1];
    if (threadIdx.x == 0) {  // last entry of this row: upper bound in the sorted key list
      int64_t lo = t_end, hi = n;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] <= key) lo = mid + 1; else hi = mid;
      }
      s_last_tile = (lo - 1) / kTile;
    }
    __syncthreads();
    const int64_t K = s_last_tile - t;  // carry-in partials: tiles t+1 .. t+K
    const int64_t m = (K + kSpanBlocks - 1) / kSpanBlocks;
    const int64_t u0 = t + 1 + blk * m;
    const int64_t u1 = min(t + 1 + K, u0 + m);
    if (u0 < u1 && lane_on) {
      float4 acc = ldg4(carry_in + u0 * D + lane * 4);
#pragma unroll 8
      for (int64_t u = u0 + 1; u < u1; ++u) add4(acc, ldg4(carry_in + u * D + lane * 4));
      *reinterpret_cast<float4*>(s_part + blk * D + lane * 4) = acc;
    }
    __syncthreads();
    if (blk == 0 && lane_on) {
      float4 acc = ldg4(carry_out + t * D + lane * 4);
      const int n_blk = static_cast<int>((K + m - 1) / m);
      for (int j = 0; j < n_blk; ++j) add4(acc, *reinterpret_cast<const float4*>(s_part + j * D + lane * 4));
      *reinterpret_cast<float4*>(d_table + static_cast<int64_t>(key) * D + lane * 4) = acc;
    }
    __syncthreads();
  }
}

int lanes_per_row(int D) {
  int lpr = 1;
  while (lpr * 4 < D) lpr <<= 1;
  return lpr;
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
  float* carry_in;
  float* carry_out;
  int* span_count;
  int* span_tiles;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

int carve_scatter_workspace(void* base, int64_t n, int D, ScatterWorkspace* w) {
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<unsigned*>(nullptr),
                                                  static_cast<unsigned*>(nullptr), static_cast<int*>(nullptr),
                                                  static_cast<int*>(nullptr), n > 0 ? n : 1, 0, 32,
                                                  static_cast<cudaStream_t>(nullptr));
  if (e != cudaSuccess) return fail(AREAD_ERR_CUDA, "cub size query failed: %s", cudaGetErrorString(e));
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
  w->carry_in = reinterpret_cast<float*>(take(static_cast<size_t>(n_tiles + 1) * D * 4));
  w->carry_out = reinterpret_cast<float*>(take(static_cast<size_t>(n_tiles + 1) * D * 4));
  w->span_count = reinterpret_cast<int*>(take(256));
  w->span_tiles = reinterpret_cast<int*>(take(static_cast<size_t>(n_tiles + 1) * 4));
  w->cub_temp = take(cub_bytes);
  w->cub_bytes = cub_bytes;
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
  const int64_t n_slots = a.batch * a.plan.n_fields;
  const int64_t groups = (n_slots + kUnroll - 1) / kUnroll;
  int64_t grid = (groups * LPR + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;  // 8 CTAs of 256 threads fill an SM
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  AREAD_LAUNCH((gather_kernel<LPR, kUnroll>), static_cast<unsigned>(grid), kThreads, plan_smem_bytes(a.plan), stream,
               a);
  return AREAD_OK;
}

template <int LPR>
int launch_scatter(const aread_scatter_args& a, const ScatterWorkspace& w, int64_t n, cudaStream_t stream) {
  const int D = a.plan.embed_dim;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int groups_per_cta = kThreads / LPR;
  const size_t smem = sizeof(int) * 2 * static_cast<size_t>(a.plan.n_cols);
  AREAD_LAUNCH((scatter_tile_kernel<LPR>), static_cast<unsigned>((n_tiles + groups_per_cta - 1) / groups_per_cta),
               kThreads, smem, stream, a.plan, w.keys_out, w.pos_out, n, n_tiles, a.d_out, a.d_table, w.carry_in,
               w.carry_out, w.span_count, w.span_tiles);
  AREAD_LAUNCH((scatter_span_kernel<LPR>), kNumSMs * 2, kSpanBlocks * LPR, sizeof(float) * kSpanBlocks * D, stream, D,
               w.keys_out, n, w.carry_in, w.carry_out, w.span_count, w.span_tiles, a.d_table);
  return AREAD_OK;
}

}  // namespace
}  // namespace aread

extern "C" {

int aread_gather_fwd(const aread_gather_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "gather: null args");
  const aread_gather_args& a = *args;
  if (int rc = check_plan(a.plan)) return rc;
  AREAD_REQUIRE(a.batch >= 0, "gather: negative batch");
  if (a.batch == 0) return AREAD_OK;
  AREAD_REQUIRE(a.x && a.table && a.out && a.status, "gather: null pointer");
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(a.table) | reinterpret_cast<uintptr_t>(a.out)) % 16 == 0,
                "gather: table/out must be 16-byte aligned");
  AREAD_REQUIRE(a.out_bf16 == nullptr || reinterpret_cast<uintptr_t>(a.out_bf16) % 8 == 0,
                "gather: out_bf16 must be 8-byte aligned");
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
  AREAD_REQUIRE(a.d_table != nullptr, "scatter: null d_table");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D = a.plan.embed_dim;
  if (a.zero_fill)
    AREAD_CUDA(cudaMemsetAsync(a.d_table, 0, static_cast<size_t>(a.plan.n_rows) * D * sizeof(float), stream));
  const int64_t n = a.batch * a.plan.n_cols;
  if (n == 0) return AREAD_OK;
  AREAD_REQUIRE(n < (int64_t{1} << 31), "scatter: %lld lookups exceed int32 positions", static_cast<long long>(n));
  AREAD_REQUIRE(a.x && a.d_out && a.workspace, "scatter: null pointer");
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(a.d_table) | reinterpret_cast<uintptr_t>(a.d_out) |
                 reinterpret_cast<uintptr_t>(a.workspace)) % 16 == 0,
                "scatter: d_table/d_out/workspace must be 16-byte aligned");
  ScatterWorkspace w;
  if (int rc = carve_scatter_workspace(a.workspace, n, D, &w)) return rc;
  if (w.total > a.workspace_bytes)
    return fail(AREAD_ERR_WORKSPACE, "scatter: workspace %zu < %zu bytes", a.workspace_bytes, w.total);

  AREAD_CUDA(cudaMemsetAsync(w.span_count, 0, sizeof(int), stream));
  {
    int64_t grid = (n + kThreads - 1) / kThreads;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    AREAD_LAUNCH(scatter_keys_kernel, static_cast<unsigned>(grid), kThreads, 0, stream, a.plan, a.x, n, w.keys_in,
                 w.pos_in);
  }
  // Stable LSD radix sort by table row (CUB, the CUDA toolkit's header library); the payload is the
  // flattened (sample, column) position, so equal rows keep the reference's accumulation order.
  size_t cub_bytes = w.cub_bytes;
  AREAD_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, cub_bytes, w.keys_in, w.keys_out, w.pos_in, w.pos_out, n, 0,
                                             key_bits(a.plan.n_rows), stream));
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  int rc;
  switch (lanes_per_row(D)) {
    case 1: rc = launch_scatter<1>(a, w, n, stream); break;
    case 2: rc = launch_scatter<2>(a, w, n, stream); break;
    case 4: rc = launch_scatter<4>(a, w, n, stream); break;
    case 8: rc = launch_scatter<8>(a, w, n, stream); break;
    case 16: rc = launch_scatter<16>(a, w, n, stream); break;
    default: rc = launch_scatter<32>(a, w, n, stream); break;
  }
  if (rc) return rc;
  if (a.sorted_rows)
    AREAD_CUDA(cudaMemcpyAsync(a.sorted_rows, w.keys_out, n * 4, cudaMemcpyDeviceToDevice, stream));
  if (a.sorted_pos)
    AREAD_CUDA(cudaMemcpyAsync(a.sorted_pos, w.pos_out, n * 4, cudaMemcpyDeviceToDevice, stream));
  return AREAD_OK;
}

}  // extern "C"
