// HEI tower layers (Linear -> BatchNorm1d -> ReLU -> Dropout, model/layer.py:221-229; the towers that run under the
// current HEMP mask, model/aread.py:297-321) on the 5th-generation tensor cores with fp32-grade arithmetic.
//
// The towers of a level are 8..64 wide -- far below a UMMA tile -- so a level's ACTIVE towers are packed
// block-diagonally: 64 consecutive input columns (64 / k towers) form one block whose [64 x NB] weight matrix
// (NB = 64 n / k output columns) is zero outside the towers' own k x n squares.  A CTA owns one block and walks
// 128-row units of it; the activations stay fp32 in HBM and never exist as bf16 there.
//
// One warp issues the MMAs; the other 16 are identical WORKERS that do both sides of the tensor-core step, so the
// per-row arithmetic (which a single warp executes at ~0.25 instructions per cycle) is spread over all of them:
//   produce(u)    a worker reads its 8 rows of unit u (coalesced 128-bit loads), applies the previous layer's
//                 BatchNorm / ReLU / Dropout on the fly, splits every value into bf16 hi + lo and writes both halves
//                 into 128B-swizzled shared-memory tiles (the layout TMA would have produced)
//   MMA(u)        tcgen05.mma, three products per tile: hi.hi + lo.hi + hi.lo, fp32 accumulate in TMEM
//   drain(u - 2)  a worker takes a [32 rows x 16 columns] slice of the accumulator (tcgen05.ld), adds the bias, writes
//                 it to HBM through a small shared-memory buffer (whole 64-byte row segments) and accumulates the
//                 column sums of its slice in registers (BatchNorm statistics forward, the BatchNorm-backward sums of
//                 the layer below backward); the sums of all workers are combined once, in a fixed order
// The loads of unit u are issued before drain(u - 2) and consumed after it: they are in flight while the worker drains.
//
// Backward: the dz tile written once to shared memory is BOTH the K-major A operand of the input gradient
// dz . W and the MN-major B operand of the weight gradient x^T . dz (rows = reduction dimension), whose A operand
// [x_hi | x_lo] is the forward input recomputed by the workers; the weight gradient accumulates in TMEM over all
// units of the CTA and leaves once, as one partial per CTA (summed in CTA order: bit-reproducible).
//
// HBM traffic per layer: input + output once (forward), z + d_out + input + d_in once (backward) -- the same as
// the CUDA-core kernels of hei.cu, which remain the path for shapes outside the block packing.
#include <cstdlib>

#include "bn_common.cuh"
#include "common.cuh"
#include "hei_tc.cuh"
#include "sm100_ptx.cuh"
#include "tensor_map.cuh"

namespace aread {
namespace {

constexpr int kTile = 128;          // rows per unit (UMMA M)
constexpr int kBlk = 64;            // input columns per block: one 128-byte swizzle row of bf16
constexpr int kTileBytes = kTile * 128;   // one bf16 [128][64] operand tile
constexpr int kWorkers = 16;        // worker w: rows 8w..8w+7 of a unit; TMEM quadrant (w + 1) % 4, column slice w / 4
constexpr int kThreads = (1 + kWorkers) * 32;   // 544: 96 registers per thread
constexpr int kFwdStages = 3, kBwdStages = 2;
constexpr int kSliceBytes = 32 * 64;            // a worker's [32 rows][16 fp32] slice buffer
constexpr int kMinRows = 512;       // below this the CUDA-core kernels are as fast

struct Plan {
  int T, NB, n_blocks, n_tiles, nb_ctas, grid;
};

bool shape_ok(int64_t m, int groups, int k, int n) {
  if (m < kMinRows || groups <= 0 || groups > 1024) return false;
  if (k != 16 && k != 32 && k != 64) return false;
  if (n <= 0 || n % 4 != 0) return false;
  const int nb = (kBlk / k) * n;
  return nb == 32 || nb == 64;
}

Plan make_plan(int64_t m, int groups, int k, int n) {
  Plan p;
  p.T = kBlk / k;
  p.NB = p.T * n;
  p.n_blocks = (groups + p.T - 1) / p.T;
  p.n_tiles = static_cast<int>((m + kTile - 1) / kTile);
  int per_block = sm_count() / p.n_blocks;
  if (per_block < 1) per_block = 1;
  p.nb_ctas = p.n_tiles < per_block ? p.n_tiles : per_block;
  p.grid = p.nb_ctas * p.n_blocks;
  return p;
}

__device__ __forceinline__ uint32_t sw128(int row, int chunk) {   // byte offset of a 16-byte chunk in a swizzled tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
// a worker's slice buffer: [32 rows][4 pieces of 16 bytes], pieces permuted so that the three access patterns
// (thread = row; 8 rows x 4 pieces per instruction, twice) touch 32 banks per 8 lanes
__device__ __forceinline__ uint32_t slice_off(int row, int piece) {
  return static_cast<uint32_t>(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
}

// x -> bf16 hi + bf16 lo with hi + lo = x up to 2^-17 relative
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
    const float2 hf = __bfloat1622float2(hb);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
    h[q] = *reinterpret_cast<const uint32_t*>(&hb);
    l[q] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void store_split(uint8_t* tile_hi, uint8_t* tile_lo, int row, int chunk, const float (&x)[8]) {
  uint4 hi, lo;
  split8(x, hi, lo);
  const uint32_t off = sw128(row, chunk);
  *reinterpret_cast<uint4*>(tile_hi + off) = hi;
  *reinterpret_cast<uint4*>(tile_lo + off) = lo;
}

// thread = row: 16 fp32 of one row -> the slice buffer
__device__ __forceinline__ void write_slice(uint8_t* buf, int lane, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(buf + slice_off(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// the slice -> global rows row0.., columns col0..+15: an instruction writes eight whole 64-byte row segments
__device__ __forceinline__ void copy_out_slice(const uint8_t* buf, int lane, float* __restrict__ dst, int64_t ld,
                                               int64_t row0, int64_t m, int col0, int width) {
  const int piece = lane & 3, rsub = lane >> 2;
  const bool col_ok = col0 + piece * 4 < width;
  float4 val[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) val[i] = *reinterpret_cast<const float4*>(buf + slice_off(i * 8 + rsub, piece));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + i * 8 + rsub;
    if (col_ok && row < m) *reinterpret_cast<float4*>(dst + row * ld + col0 + piece * 4) = val[i];
  }
}
// per-lane partial column sums of the slice: the lane owns columns 4 * (lane & 3) .. + 3 and rows (lane >> 2) + 8 i
__device__ __forceinline__ void slice_moments(const uint8_t* buf, int lane, int n_rows, const float (&piv)[4],
                                              float (&s)[4], float (&q)[4]) {
  const int piece = lane & 3, rsub = lane >> 2;
  float4 val[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) val[i] = *reinterpret_cast<const float4*>(buf + slice_off(i * 8 + rsub, piece));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool live = i * 8 + rsub < n_rows;
    const float d[4] = {val[i].x - piv[0], val[i].y - piv[1], val[i].z - piv[2], val[i].w - piv[3]};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float dv = live ? d[c] : 0.f;
      s[c] += dv;
      q[c] = fmaf(dv, dv, q[c]);
    }
  }
}
__device__ __forceinline__ void slice_sums(const uint8_t* buf, int lane, float (&s)[4]) {
  const int piece = lane & 3, rsub = lane >> 2;
  float4 val[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) val[i] = *reinterpret_cast<const float4*>(buf + slice_off(i * 8 + rsub, piece));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[0] += val[i].x; s[1] += val[i].y; s[2] += val[i].z; s[3] += val[i].w;
  }
}
// The workers' per-lane partial sums part[worker][lane][8] (two quantities x four columns) -> the sum of column j of
// the block for quantity `which`: the four quadrant workers of the slice in worker order, their eight row groups in
// lane order.
__device__ __forceinline__ float combine_column(const float* part, int j, int which) {
  const int sl = j >> 4, piece = (j >> 2) & 3, c = j & 3;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* pw = part + (sl * 4 + i) * 32 * 8;
#pragma unroll
    for (int rg = 0; rg < 8; ++rg) sum += pw[(rg * 4 + piece) * 8 + which * 4 + c];
  }
  return sum;
}

// =================================================================================================== forward
struct FwdParams {
  aread_hei_layer_fwd_args a;
  int T, NB, n_blocks, n_tiles, nb_ctas;
  uint32_t thr;
  float keep_scale;
  int do_stats;
  float* partial;   // [nb_ctas][2][groups * n]
  float* pivot;     // [groups * n]: the shift of the variance sums (the running mean before this step's update)
};

struct FwdSmem {
  static constexpr int kW = 0;                                   // W_hi, W_lo: [NB <= 64 rows][128 B] each
  static constexpr int kStage = 2 * 64 * 128;                    // 16 KB
  static constexpr int kStageBytes = 2 * kTileBytes;             // A_hi, A_lo
  static constexpr int kSlice = kStage + kFwdStages * kStageBytes;
  static constexpr int kSmall = kSlice + kWorkers * kSliceBytes; // pivot[64], bias[64], sc[64], sh[64]
  static constexpr int kSmallBytes = 4 * 64 * 4;
  static constexpr int kBar = kSmall + kSmallBytes;
  static constexpr int kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kThreads, 1) hei_tc_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* empty_bar = full_bar + kFwdStages;
  uint64_t* acc_full = empty_bar + kFwdStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_pivot = reinterpret_cast<float*>(smem + FwdSmem::kSmall);
  float* s_bias = s_pivot + 64;
  float* s_sc = s_bias + 64;      // BatchNorm scale / shift of the layer below per input column of the block
  float* s_sh = s_sc + 64;

  const aread_hei_layer_fwd_args& a = p.a;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int K = a.k, N = a.n, G = a.groups, NB = p.NB, T = p.T;
  const int b = blockIdx.x % p.n_blocks, ci = blockIdx.x / p.n_blocks;
  const int src_width = G * K, out_width = G * N;
  const int n_units = (p.n_tiles - ci + p.nb_ctas - 1) / p.nb_ctas;
  const uint64_t seed = seed_of(a);
  constexpr uint32_t kTmemCols = 128;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFwdStages; ++s) {
      ptx::mbar_init(&full_bar[s], kWorkers);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kWorkers);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, kTmemCols);

  // the block's weights: row j = (tower, output column), 64 reduction columns of which the tower's own k are nonzero
  for (int idx = threadIdx.x; idx < NB * 8; idx += kThreads) {
    const int j = idx >> 3, c = idx & 7;
    const int tl = j / N, jj = j - tl * N, g = b * T + tl;
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = c * 8 + e;
      w[e] = (g < G && i / K == tl) ? __ldg(a.weight + (static_cast<int64_t>(g) * N + jj) * K + (i - tl * K)) : 0.f;
    }
    store_split(smem + FwdSmem::kW, smem + FwdSmem::kW + 64 * 128, j, c, w);
  }
  if (threadIdx.x >= kThreads - 64) {   // per-column constants
    const int j = threadIdx.x - (kThreads - 64);
    const int col = b * NB + j;
    float bias = 0.f, piv = 0.f;
    if (j < NB && col < out_width) {
      bias = a.bias ? __ldg(a.bias + col) : 0.f;
      if (p.do_stats) {
        piv = a.running_mean[col];
        if (ci == 0) p.pivot[col] = piv;
      }
    }
    s_bias[j] = bias;
    s_pivot[j] = piv;
    const int scol = b * kBlk + j;
    const bool ok = a.src_scale != nullptr && scol < src_width;
    s_sc[j] = ok ? __ldg(a.src_scale + scol) : 1.f;
    s_sh[j] = ok ? __ldg(a.src_shift + scol) : 0.f;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = ptx::umma_idesc_bf16_ab(kTile, NB, false, false);
      const uint32_t w_hi = ptx::smem_u32(smem + FwdSmem::kW), w_lo = w_hi + 64 * 128;
      for (int u = 0; u < n_units; ++u) {
        const int stage = u % kFwdStages, acc = u & 1;
        ptx::mbar_wait(&acc_empty[acc], ((u >> 1) & 1) ^ 1);
        ptx::mbar_wait(&full_bar[stage], (u / kFwdStages) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t a_hi = ptx::smem_u32(smem + FwdSmem::kStage + stage * FwdSmem::kStageBytes), a_lo = a_hi + kTileBytes;
        const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
          const uint32_t sa = seg == 1 ? a_lo : a_hi, sb = seg == 2 ? w_lo : w_hi;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_bf16(d_tmem, ptx::umma_desc_k_sw128(sa + kk * 32, 1024), ptx::umma_desc_k_sw128(sb + kk * 32, 1024),
                           idesc, (seg | kk) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
        ptx::umma_commit(&acc_full[acc]);
      }
    }
  } else {  // ===== workers =====
    const int w = warp - 1;
    const int quad = warp % 4, sl = w / 4;
    const bool has_slice = sl * 16 < NB;
    uint8_t* buf = smem + FwdSmem::kSlice + w * kSliceBytes;
    // produce role: rows 8w + (lane >> 3) and + 4, 8 columns
    const int chunk = lane & 7, r0 = w * 8 + (lane >> 3);
    const int col0 = b * kBlk + chunk * 8;
    const bool col_ok = col0 < src_width;          // k is a multiple of 8: a chunk is inside the tensor or outside
    const bool src_bn = a.src_scale != nullptr;
    // drain role: statistics of columns sl * 16 + 4 * (lane & 3) .. + 3 over the rows (lane >> 2) + 8 i
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f}, piv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) piv[c] = s_pivot[(sl * 16 + (lane & 3) * 4 + c) & 63];

    auto issue = [&](float4 (&raw)[2][2], int uu) {
      const int64_t tile_row = (static_cast<int64_t>(ci) + static_cast<int64_t>(uu) * p.nb_ctas) * kTile;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int64_t row = tile_row + r0 + 4 * it;
        if (uu < n_units && col_ok && row < a.m) {
          const float* sp = a.src + row * a.ld_src + col0;
          raw[it][0] = ldg4(sp);
          raw[it][1] = ldg4(sp + 4);
        } else {
          raw[it][0] = raw[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto drain = [&](int ud) {
      const int acc = ud & 1;
      ptx::mbar_wait(&acc_full[acc], (ud >> 1) & 1);
      ptx::tc_fence_after_sync();
      float v[16];
      if (has_slice) ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64 + sl * 16, v);
      ptx::tc_fence_before_sync();
      __syncwarp();                                  // also: the previous unit's reads of `buf` are done
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      if (has_slice) {
        const int64_t row0 = (static_cast<int64_t>(ci) + static_cast<int64_t>(ud) * p.nb_ctas) * kTile + quad * 32;
        const int n_rows = a.m - row0 >= 32 ? 32 : (a.m > row0 ? static_cast<int>(a.m - row0) : 0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + sl * 16 + j4 * 4);
          v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
        }
        write_slice(buf, lane, v);
        __syncwarp();
        copy_out_slice(buf, lane, a.z, out_width, row0, a.m, b * NB + sl * 16, out_width);
        if (p.do_stats) slice_moments(buf, lane, n_rows, piv, cs, cq);
      }
    };
    auto convert = [&](const float4 (&raw)[2][2], int uu) {
      const int64_t tile_row = (static_cast<int64_t>(ci) + static_cast<int64_t>(uu) * p.nb_ctas) * kTile;
      const int stage = uu % kFwdStages;
      ptx::mbar_wait(&empty_bar[stage], ((uu / kFwdStages) & 1) ^ 1);
      uint8_t* a_hi = smem + FwdSmem::kStage + stage * FwdSmem::kStageBytes;
      uint8_t* a_lo = a_hi + kTileBytes;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = r0 + 4 * it;
        const int64_t row = tile_row + r;
        float x[8] = {raw[it][0].x, raw[it][0].y, raw[it][0].z, raw[it][0].w,
                      raw[it][1].x, raw[it][1].y, raw[it][1].z, raw[it][1].w};
        if (src_bn) {
          const bool live = col_ok && row < a.m;
          const uint64_t flat = static_cast<uint64_t>(row) * src_width + col0;
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const float4 f_sc = *reinterpret_cast<const float4*>(s_sc + chunk * 8 + q4 * 4);
            const float4 f_sh = *reinterpret_cast<const float4*>(s_sh + chunk * 8 + q4 * 4);
            const float sc[4] = {f_sc.x, f_sc.y, f_sc.z, f_sc.w}, sh[4] = {f_sh.x, f_sh.y, f_sh.z, f_sh.w};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t h = 0xffffffffu;
              if (p.thr != 0u) h = dropout_pair_hash(seed, a.src_salt, (flat >> 1) + q4 * 2 + q);
              const bool k0 = (h & 0xffffu) >= p.thr, k1 = (h >> 16) >= p.thr;
              const int e = q4 * 4 + q * 2;
              x[e] = live ? act_value(x[e], sc[q * 2], sh[q * 2], k0, p.keep_scale) : 0.f;
              x[e + 1] = live ? act_value(x[e + 1], sc[q * 2 + 1], sh[q * 2 + 1], k1, p.keep_scale) : 0.f;
            }
          }
        }
        store_split(a_hi, a_lo, r, chunk, x);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
    };

    // the loads of unit u are issued before drain(u - 2) and consumed after it (two register sets in flight were
    // measured slower: the kernel is bound by instruction issue, not by the latency of these loads)
    for (int u = 0; u < n_units + 2; ++u) {
      float4 raw[2][2];
      issue(raw, u);
      if (u >= 2) drain(u - 2);
      if (u < n_units) convert(raw, u);
    }
    if (p.do_stats) {   // per-lane partial sums -> one partial per CTA; the tiles are dead by now
      float* part = reinterpret_cast<float*>(smem + FwdSmem::kStage);
      float* mine = part + (w * 32 + lane) * 8;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mine[c] = cs[c];
        mine[4 + c] = cq[c];
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const int tt = w * 32 + lane;
      if (tt < 2 * NB) {
        const int which = tt / NB, j = tt - which * NB;
        const int col = b * NB + j;
        if (col < out_width) p.partial[(static_cast<int64_t>(ci) * 2 + which) * out_width + col] = combine_column(part, j, which);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =================================================================================================== backward
struct BwdParams {
  aread_hei_layer_bwd_args a;
  int T, n_blocks, n_tiles, nb_ctas;
  uint32_t thr, src_thr;
  float keep_scale, src_keep_scale;
  float* partial_w;   // [nb_ctas][groups * n * k]
  float* partial_s;   // [nb_ctas][2][groups * k]
};

struct BwdSmem {
  static constexpr int kW = 0;                                   // Wt_hi, Wt_lo: [64 rows][128 B] each
  static constexpr int kStage = 2 * 64 * 128;
  static constexpr int kStageBytes = 4 * kTileBytes;             // dz_hi, dz_lo, x_hi, x_lo
  static constexpr int kSlice = kStage + kBwdStages * kStageBytes;   // two slice buffers per worker
  static constexpr int kSmall = kSlice + kWorkers * 2 * kSliceBytes; // 8 parameter rows of 64
  static constexpr int kSmallBytes = 8 * 64 * 4;
  static constexpr int kBar = kSmall + kSmallBytes;
  static constexpr int kTotal = kBar + 256 + 1024;
  static constexpr int kPartOffset = 80 * 1024;                  // per-lane sums, behind the dW scratch (final phase)
};

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) hei_tc_bwd_kernel(const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* empty_bar = full_bar + kBwdStages;
  uint64_t* acc_full = empty_bar + kBwdStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_done = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_done + 1);
  float* s_sc = reinterpret_cast<float*>(smem + BwdSmem::kSmall);   // produce: src scale / shift per input column
  float* s_sh = s_sc + 64;
  float* s_ca = s_sh + 64;      // drain: xhat = x * ca - cb for the elements the forward kept (x > 0)
  float* s_cb = s_ca + 64;
  float* s_pa = s_cb + 64;      // dz = pa * dy - pc * z + pd, y = pa * z + psh per output column
  float* s_psh = s_pa + 64;
  float* s_pc = s_psh + 64;
  float* s_pd = s_pc + 64;

  const aread_hei_layer_bwd_args& a = p.a;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int K = a.k, N = a.n, G = a.groups, T = p.T;
  const int b = blockIdx.x % p.n_blocks, ci = blockIdx.x / p.n_blocks;
  const int src_width = G * K, out_width = G * N;
  const int n_units = (p.n_tiles - ci + p.nb_ctas - 1) / p.nb_ctas;
  const bool src_bn = a.src_scale != nullptr;
  const uint64_t seed = seed_of(a);
  constexpr uint32_t kTmemCols = 256;         // [0,128): two d_in accumulators; [128,128+NB), [192,192+NB): dW
  constexpr int kDzChunks = NB / 8;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      ptx::mbar_init(&full_bar[s], kWorkers);
      ptx::mbar_init(&empty_bar[s], src_bn ? 1 + kWorkers : 1);   // the drain reads x from the stage when src_bn
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kWorkers);
    }
    ptx::mbar_init(w_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, kTmemCols);

  // W^T of the block: row i = input column, reduction index j = (tower, output column)
  for (int idx = threadIdx.x; idx < 64 * kDzChunks; idx += kThreads) {
    const int i = idx / kDzChunks, c = idx - i * kDzChunks;
    const int tl = i / K, ii = i - tl * K, g = b * T + tl;
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = c * 8 + e;
      w[e] = (g < G && j / N == tl) ? __ldg(a.weight + (static_cast<int64_t>(g) * N + (j - tl * N)) * K + ii) : 0.f;
    }
    store_split(smem + BwdSmem::kW, smem + BwdSmem::kW + 64 * 128, i, c, w);
  }
  if (threadIdx.x >= kThreads - 64) {     // per-column constants
    const int j = threadIdx.x - (kThreads - 64);
    {
      // y = z * scale + shift, xhat = (z - mean) * rstd  =>  for a kept element x = y * keep_scale > 0:
      // xhat = x * rstd / (scale * keep_scale) - (shift / scale + mean) * rstd
      const int col = b * kBlk + j;
      const bool ok = src_bn && col < src_width;
      const float sc = ok ? __ldg(a.src_scale + col) : 1.f;
      const float sh = ok ? __ldg(a.src_shift + col) : 0.f;
      const float rs = ok ? __ldg(a.src_rstd + col) : 0.f;
      const float mu = ok ? __ldg(a.src_mean + col) : 0.f;
      const float inv = sc != 0.f ? 1.f / sc : 0.f;
      s_sc[j] = sc;
      s_sh[j] = sh;
      s_ca[j] = rs * inv / p.src_keep_scale;
      s_cb[j] = fmaf(sh, inv, mu) * rs;
    }
    {
      // dz = A * dy - B - (z - mean) * C, A = scale, B = scale * mean(dy), C = scale * rstd * mean(dy * xhat);
      // folded: dz = A * dy - C * z + (mean * C - B)
      const int col = b * NB + j;
      const bool ok = j < NB && col < out_width;
      const float sc = ok ? __ldg(a.scale + col) : 0.f;
      float pc = 0.f, pd = 0.f;
      if (ok && !a.bn_skip) {
        pc = sc * __ldg(a.rstd + col) * __ldg(a.coef + out_width + col);
        pd = fmaf(__ldg(a.mean + col), pc, -sc * __ldg(a.coef + col));
      }
      s_pa[j] = sc;
      s_psh[j] = ok ? __ldg(a.shift + col) : 0.f;
      s_pc[j] = pc;
      s_pd[j] = pd;
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc_d = ptx::umma_idesc_bf16_ab(kTile, 64, false, false);
      constexpr uint32_t idesc_w = ptx::umma_idesc_bf16_ab(kTile, NB, true, true);
      const uint32_t w_hi = ptx::smem_u32(smem + BwdSmem::kW), w_lo = w_hi + 64 * 128;
      for (int u = 0; u < n_units; ++u) {
        const int stage = u % kBwdStages, acc = u & 1;
        ptx::mbar_wait(&acc_empty[acc], ((u >> 1) & 1) ^ 1);
        ptx::mbar_wait(&full_bar[stage], (u / kBwdStages) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t dz_hi = ptx::smem_u32(smem + BwdSmem::kStage + stage * BwdSmem::kStageBytes);
        const uint32_t dz_lo = dz_hi + kTileBytes, x_hi = dz_lo + kTileBytes;
        // d_in = dz . W : K-major dz [128 rows][NB], K-major W^T [64][NB]
        const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
          const uint32_t sa = seg == 1 ? dz_lo : dz_hi, sb = seg == 2 ? w_lo : w_hi;
#pragma unroll
          for (int kk = 0; kk < NB / 16; ++kk)
            ptx::umma_bf16(d_tmem, ptx::umma_desc_k_sw128(sa + kk * 32, 1024), ptx::umma_desc_k_sw128(sb + kk * 32, 1024),
                           idesc_d, (seg | kk) != 0);
        }
        ptx::umma_commit(&acc_full[acc]);
        // dW^T (+)= [x_hi | x_lo]^T . dz_hi  and  [x_hi | x_lo]^T . dz_lo : both operands MN-major, K = the 128 rows
#pragma unroll
        for (int kk = 0; kk < kTile / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_mn_sw128(x_hi + kk * 2048, kTileBytes, 1024);
          ptx::umma_bf16(tmem_base + 128, da, ptx::umma_desc_mn_sw128(dz_hi + kk * 2048, kTileBytes, 1024), idesc_w,
                         u != 0 || kk != 0);
          ptx::umma_bf16(tmem_base + 192, da, ptx::umma_desc_mn_sw128(dz_lo + kk * 2048, kTileBytes, 1024), idesc_w,
                         u != 0 || kk != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
      }
      ptx::umma_commit(w_done);
    }
  } else {  // ===== workers =====
    const int w = warp - 1;
    const int quad = warp % 4, sl = w / 4;
    uint8_t* buf_a = smem + BwdSmem::kSlice + w * 2 * kSliceBytes;
    uint8_t* buf_b = buf_a + kSliceBytes;
    // produce role: dz of rows 8w.. (a fixed chunk of 8 output columns per lane) and x (8 input columns per lane)
    constexpr int kDzRows = 32 / kDzChunks;            // rows a warp covers per pass: 4 (NB = 64) or 8 (NB = 32)
    constexpr int kDzPasses = 8 / kDzRows;
    const int dc = lane % kDzChunks, dsub = lane / kDzChunks;
    const int dcol0 = b * NB + dc * 8;
    const bool dcol_ok = dcol0 < out_width;
    const int xc = lane & 7, xsub = lane >> 3;
    const int xcol0 = b * kBlk + xc * 8;
    const bool xcol_ok = xcol0 < src_width;
    // drain role: BatchNorm-backward sums of input columns sl * 16 + 4 * (lane & 3) .. + 3
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};

    // A unit's rows are loaded in two halves, each into its own register set, and a set is refilled with the next
    // unit's rows as soon as it has been converted: half a unit to a whole unit of loads is in flight per worker while
    // it converts the other half and drains.
    struct HalfRegs {
      float4 z[2], d[2], x[2];
    };
    auto issue = [&](HalfRegs& h, int uu, int hf) {
      const int64_t row_base = (static_cast<int64_t>(ci) + static_cast<int64_t>(uu) * p.nb_ctas) * kTile + w * 8;
      const bool on = uu < n_units;
      if (hf < kDzPasses) {
        const int64_t row = row_base + hf * kDzRows + dsub;
        if (on && dcol_ok && row < a.m) {
          const float* zp = a.z + row * out_width + dcol0;
          const float* dp = a.d_out + row * out_width + dcol0;
          h.z[0] = ldg4(zp); h.z[1] = ldg4(zp + 4);
          h.d[0] = ldg4(dp); h.d[1] = ldg4(dp + 4);
        } else {
          h.z[0] = h.z[1] = h.d[0] = h.d[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const int64_t row = row_base + hf * 4 + xsub;
      if (on && xcol_ok && row < a.m) {
        const float* sp = a.src + row * a.ld_src + xcol0;
        h.x[0] = ldg4(sp); h.x[1] = ldg4(sp + 4);
      } else {
        h.x[0] = h.x[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto convert = [&](const HalfRegs& h, int uu, int hf) {
      const int64_t row_base = (static_cast<int64_t>(ci) + static_cast<int64_t>(uu) * p.nb_ctas) * kTile + w * 8;
      const int stage = uu % kBwdStages;
      uint8_t* dz_hi = smem + BwdSmem::kStage + stage * BwdSmem::kStageBytes;
      uint8_t* dz_lo = dz_hi + kTileBytes;
      uint8_t* x_hi = dz_lo + kTileBytes;
      uint8_t* x_lo = x_hi + kTileBytes;
      if (hf < kDzPasses) {   // dy = d_out * [y > 0] * keep / (1 - p); dz = pa * dy - pc * z + pd
        const int r = w * 8 + hf * kDzRows + dsub;
        const int64_t row = row_base + hf * kDzRows + dsub;
        const bool live = dcol_ok && row < a.m;
        const float z[8] = {h.z[0].x, h.z[0].y, h.z[0].z, h.z[0].w, h.z[1].x, h.z[1].y, h.z[1].z, h.z[1].w};
        const float d[8] = {h.d[0].x, h.d[0].y, h.d[0].z, h.d[0].w, h.d[1].x, h.d[1].y, h.d[1].z, h.d[1].w};
        const uint64_t flat = static_cast<uint64_t>(row) * out_width + dcol0;
        float dz[8];
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const float4 f_pa = *reinterpret_cast<const float4*>(s_pa + dc * 8 + q4 * 4);
          const float4 f_sh = *reinterpret_cast<const float4*>(s_psh + dc * 8 + q4 * 4);
          const float4 f_pc = *reinterpret_cast<const float4*>(s_pc + dc * 8 + q4 * 4);
          const float4 f_pd = *reinterpret_cast<const float4*>(s_pd + dc * 8 + q4 * 4);
          const float pa[4] = {f_pa.x, f_pa.y, f_pa.z, f_pa.w}, psh[4] = {f_sh.x, f_sh.y, f_sh.z, f_sh.w};
          const float pc[4] = {f_pc.x, f_pc.y, f_pc.z, f_pc.w}, pd[4] = {f_pd.x, f_pd.y, f_pd.z, f_pd.w};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t hh = 0xffffffffu;
            if (p.thr != 0u) hh = dropout_pair_hash(seed, a.salt, (flat >> 1) + q4 * 2 + q);
            const bool kp[2] = {(hh & 0xffffu) >= p.thr, (hh >> 16) >= p.thr};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e4 = q * 2 + e2, e = q4 * 4 + e4;
              const float y = fmaf(z[e], pa[e4], psh[e4]);
              const float dyv = (y > 0.f && kp[e2]) ? d[e] * p.keep_scale : 0.f;
              dz[e] = live ? fmaf(pa[e4], dyv, fmaf(-pc[e4], z[e], pd[e4])) : 0.f;
            }
          }
        }
        store_split(dz_hi, dz_lo, r, dc, dz);
      }
      {   // x = the layer input as the forward saw it: plain, or dropout(relu(bn(z_prev))) recomputed
        const int r = w * 8 + hf * 4 + xsub;
        const int64_t row = row_base + hf * 4 + xsub;
        float x[8] = {h.x[0].x, h.x[0].y, h.x[0].z, h.x[0].w, h.x[1].x, h.x[1].y, h.x[1].z, h.x[1].w};
        if (src_bn) {
          const bool live = xcol_ok && row < a.m;
          const uint64_t flat = static_cast<uint64_t>(row) * src_width + xcol0;
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const float4 f_sc = *reinterpret_cast<const float4*>(s_sc + xc * 8 + q4 * 4);
            const float4 f_sh = *reinterpret_cast<const float4*>(s_sh + xc * 8 + q4 * 4);
            const float qs[4] = {f_sc.x, f_sc.y, f_sc.z, f_sc.w}, qh[4] = {f_sh.x, f_sh.y, f_sh.z, f_sh.w};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t hh = 0xffffffffu;
              if (p.src_thr != 0u) hh = dropout_pair_hash(seed, a.src_salt, (flat >> 1) + q4 * 2 + q);
              const bool k0 = (hh & 0xffffu) >= p.src_thr, k1 = (hh >> 16) >= p.src_thr;
              const int e = q4 * 4 + q * 2;
              x[e] = live ? act_value(x[e], qs[q * 2], qh[q * 2], k0, p.src_keep_scale) : 0.f;
              x[e + 1] = live ? act_value(x[e + 1], qs[q * 2 + 1], qh[q * 2 + 1], k1, p.src_keep_scale) : 0.f;
            }
          }
        }
        store_split(x_hi, x_lo, r, xc, x);
      }
    };
    auto drain = [&](int ud) {   // d_in slice [32 rows x 16 columns] of unit ud
      const int acc = ud & 1, dstage = ud % kBwdStages;
      ptx::mbar_wait(&acc_full[acc], (ud >> 1) & 1);
      ptx::tc_fence_after_sync();
      float v[16];
      ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 64 + sl * 16, v);
      ptx::tc_fence_before_sync();
      __syncwarp();                                  // also: the previous unit's reads of the slice buffers are done
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      const int64_t row0 = (static_cast<int64_t>(ci) + static_cast<int64_t>(ud) * p.nb_ctas) * kTile + quad * 32;
      const int col0 = b * kBlk + sl * 16;
      if (a.d_in != nullptr) {
        write_slice(buf_a, lane, v);
        __syncwarp();
        copy_out_slice(buf_a, lane, a.d_in, src_width, row0, a.m, col0, src_width);
      }
      if (src_bn) {   // sums of the BatchNorm backward of the layer below: dy = d_in * [x > 0] * keep / (1 - p)
        // x of this row comes from the split tiles of the stage: the elements the forward kept are the positive ones
        const uint8_t* x_hi = smem + BwdSmem::kStage + dstage * BwdSmem::kStageBytes + 2 * kTileBytes;
        const uint8_t* x_lo = x_hi + kTileBytes;
        const int r = quad * 32 + lane;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = sw128(r, sl * 2 + c);
          const uint4 hi = *reinterpret_cast<const uint4*>(x_hi + off);
          const uint4 lo = *reinterpret_cast<const uint4*>(x_lo + off);
          const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
          const float4 ca0 = *reinterpret_cast<const float4*>(s_ca + sl * 16 + c * 8);
          const float4 ca1 = *reinterpret_cast<const float4*>(s_ca + sl * 16 + c * 8 + 4);
          const float4 cb0 = *reinterpret_cast<const float4*>(s_cb + sl * 16 + c * 8);
          const float4 cb1 = *reinterpret_cast<const float4*>(s_cb + sl * 16 + c * 8 + 4);
          const float ca[8] = {ca0.x, ca0.y, ca0.z, ca0.w, ca1.x, ca1.y, ca1.z, ca1.w};
          const float cb[8] = {cb0.x, cb0.y, cb0.z, cb0.w, cb1.x, cb1.y, cb1.z, cb1.w};
          float xh[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t hh = hw[e >> 1], ll = lw[e >> 1];
            const float xv = __uint_as_float((e & 1) ? (hh & 0xffff0000u) : (hh << 16)) +
                             __uint_as_float((e & 1) ? (ll & 0xffff0000u) : (ll << 16));
            const float g = xv > 0.f ? v[c * 8 + e] * p.src_keep_scale : 0.f;
            v[c * 8 + e] = g;
            xh[e] = g * fmaf(xv, ca[e], -cb[e]);
          }
          __syncwarp();     // (first pass) the d_in copy-out has read buf_a
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            *reinterpret_cast<float4*>(buf_a + slice_off(lane, 2 * c + q)) =
                make_float4(v[c * 8 + 4 * q], v[c * 8 + 4 * q + 1], v[c * 8 + 4 * q + 2], v[c * 8 + 4 * q + 3]);
            *reinterpret_cast<float4*>(buf_b + slice_off(lane, 2 * c + q)) =
                make_float4(xh[4 * q], xh[4 * q + 1], xh[4 * q + 2], xh[4 * q + 3]);
          }
        }
        __syncwarp();                                          // x tile read, slices written
        if (lane == 0) ptx::mbar_arrive(&empty_bar[dstage]);   // the stage may be refilled
        slice_sums(buf_a, lane, s1);
        slice_sums(buf_b, lane, s2);
      }
    };

    HalfRegs h0, h1;
    issue(h0, 0, 0);
    issue(h1, 0, 1);
    for (int u = 0; u < n_units + 2; ++u) {
      if (u >= 2) drain(u - 2);
      if (u < n_units) {
        const int stage = u % kBwdStages;
        ptx::mbar_wait(&empty_bar[stage], ((u / kBwdStages) & 1) ^ 1);
        convert(h0, u, 0);
        issue(h0, u + 1, 0);
        convert(h1, u, 1);
        issue(h1, u + 1, 1);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
      }
    }

    // ---- final phase: weight gradient of the block and the column sums.  The scratch below reuses the tiles: every
    // worker must have finished its last drain (which reads x from them) first.
    asm volatile("bar.sync 1, 512;" ::: "memory");
    ptx::mbar_wait(w_done, 0);
    ptx::tc_fence_after_sync();
    float* scratch = reinterpret_cast<float*>(smem + BwdSmem::kStage);     // [2][128][NB + 1]
    float* part = reinterpret_cast<float*>(smem + BwdSmem::kStage + BwdSmem::kPartOffset);
    constexpr int kPitch = NB + 1;
    if (sl * 16 < NB) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {      // x^T dz_hi, x^T dz_lo: lanes of `quad`, columns of slice `sl`
        float v[16];
        ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + 128 + which * 64 + sl * 16, v);
        float* dst = scratch + (which * kTile + quad * 32 + lane) * kPitch + sl * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = v[j];
      }
    }
    {
      float* mine = part + (w * 32 + lane) * 8;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mine[c] = s1[c];
        mine[4 + c] = s2[c];
      }
    }
    ptx::tc_fence_before_sync();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int et = w * 32 + lane;      // 0..511
    for (int idx = et; idx < 64 * NB; idx += kWorkers * 32) {
      const int i = idx / NB, j = idx - i * NB;
      const int tl = i / K, g = b * T + tl;
      if (g < G && j / N == tl) {
        const float* s0 = scratch + i * kPitch + j;
        const float* s1p = scratch + (kTile + i) * kPitch + j;
        const float val = ((s0[0] + s0[64 * kPitch]) + s1p[0]) + s1p[64 * kPitch];
        p.partial_w[static_cast<int64_t>(ci) * G * N * K + (static_cast<int64_t>(g) * N + (j - tl * N)) * K + (i - tl * K)] = val;
      }
    }
    if (src_bn && et < 128) {
      const int which = et / 64, j = et % 64;
      const int col = b * kBlk + j;
      if (col < src_width) p.partial_s[(static_cast<int64_t>(ci) * 2 + which) * src_width + col] = combine_column(part, j, which);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// partials of the CTAs in CTA order; eight independent loads in flight per thread
__global__ void __launch_bounds__(256) hei_tc_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial,
                                                                  int64_t n, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  int q = 0;
  for (; q + 8 <= n_partial; q += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = partial[static_cast<int64_t>(q + u) * n + i];
#pragma unroll
    for (int u = 0; u < 8; ++u) v += t[u];
  }
  for (; q < n_partial; ++q) v += partial[static_cast<int64_t>(q) * n + i];
  out[i] = v;
}

int g_path_fwd = -1, g_path_bwd = -1;     // aread_hei_set_path overrides; -1 = the environment decides

bool env_on(const char* name) {
  const char* e = std::getenv(name);
  return e == nullptr || e[0] != '0';
}
bool tc_fwd_enabled() {
  static const bool env = env_on("AREAD_HEI_TC");
  return g_path_fwd < 0 ? env : g_path_fwd != 0;
}
bool tc_bwd_enabled() {
  static const bool env = env_on("AREAD_HEI_TC") && env_on("AREAD_HEI_TC_BWD");
  return g_path_bwd < 0 ? env : g_path_bwd != 0;
}

}  // namespace

void hei_tc_set_path(int fwd, int bwd) {
  g_path_fwd = fwd;
  g_path_bwd = bwd;
}
bool hei_tc_shape_ok(int64_t m, int groups, int k, int n) { return shape_ok(m, groups, k, n); }

bool hei_tc_usable(int64_t m, int groups, int k, int n, const float* src, int64_t ld_src) {
  return tc_fwd_enabled() && shape_ok(m, groups, k, n) && ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}
bool hei_tc_bwd_usable(int64_t m, int groups, int k, int n, const float* src, int64_t ld_src) {
  return tc_bwd_enabled() && shape_ok(m, groups, k, n) && ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}

size_t hei_tc_workspace_floats(int64_t m, int groups, int k, int n) {
  if (!shape_ok(m, groups, k, n)) return 0;
  const Plan pl = make_plan(m, groups, k, n);
  const size_t fwd = static_cast<size_t>(pl.nb_ctas) * 2 * groups * n + static_cast<size_t>(groups) * n;
  const size_t bwd = static_cast<size_t>(pl.nb_ctas) * groups * n * k + static_cast<size_t>(pl.nb_ctas) * 2 * groups * k;
  return fwd > bwd ? fwd : bwd;
}

int hei_tc_fwd(const aread_hei_layer_fwd_args& a, cudaStream_t stream) {
  const Plan pl = make_plan(a.m, a.groups, a.k, a.n);
  const int width = a.groups * a.n;
  AREAD_REQUIRE((reinterpret_cast<uintptr_t>(a.z) & 15) == 0, "hei_layer_fwd: z must be 16-byte aligned");
  FwdParams p;
  p.a = a;
  p.T = pl.T;
  p.NB = pl.NB;
  p.n_blocks = pl.n_blocks;
  p.n_tiles = pl.n_tiles;
  p.nb_ctas = pl.nb_ctas;
  const float drop = a.training ? a.src_p : 0.f;
  p.thr = drop > 0.f ? dropout_threshold(drop) : 0u;
  p.keep_scale = drop > 0.f ? 1.f / (1.f - drop) : 1.f;
  p.do_stats = (a.training && !a.bn_skip) ? 1 : 0;
  p.partial = static_cast<float*>(a.workspace);
  p.pivot = p.partial + static_cast<size_t>(pl.nb_ctas) * 2 * width;
  static uint64_t configured = 0;
  if (first_use_on_device(&configured))
    AREAD_CUDA(cudaFuncSetAttribute(hei_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal));
  AREAD_LAUNCH(hei_tc_fwd_kernel, pl.grid, kThreads, FwdSmem::kTotal, stream, p);
  aread_bn_act_args f = {};
  f.m = a.m;
  f.width = width;
  f.training = a.training;
  f.bn_skip = a.bn_skip;
  f.momentum = a.momentum;
  f.eps = a.eps;
  f.z = p.pivot;            // bn_finalize reads the pivot of column c at z[c]
  f.ldz = width;
  f.gamma = a.gamma;
  f.beta = a.beta;
  f.running_mean = a.running_mean;
  f.running_var = a.running_var;
  f.mean = a.mean;
  f.rstd = a.rstd;
  f.scale = a.scale;
  f.shift = a.shift;
  AREAD_LAUNCH(bn_finalize_kernel, ceil_div(f.width, 32), kBnThreads, 0, stream, f, p.partial, pl.nb_ctas);
  return AREAD_OK;
}

int hei_tc_bwd(const aread_hei_layer_bwd_args& a, cudaStream_t stream) {
  const Plan pl = make_plan(a.m, a.groups, a.k, a.n);
  const int src_width = a.groups * a.k;
  const bool src_bn = a.src_scale != nullptr;
  BwdParams p;
  p.a = a;
  p.T = pl.T;
  p.n_blocks = pl.n_blocks;
  p.n_tiles = pl.n_tiles;
  p.nb_ctas = pl.nb_ctas;
  p.thr = a.p > 0.f ? dropout_threshold(a.p) : 0u;
  p.keep_scale = a.p > 0.f ? 1.f / (1.f - a.p) : 1.f;
  p.src_thr = a.src_p > 0.f ? dropout_threshold(a.src_p) : 0u;
  p.src_keep_scale = a.src_p > 0.f ? 1.f / (1.f - a.src_p) : 1.f;
  const int64_t n_w = static_cast<int64_t>(a.groups) * a.n * a.k;
  p.partial_w = static_cast<float*>(a.workspace);
  p.partial_s = p.partial_w + static_cast<size_t>(pl.nb_ctas) * n_w;
  AREAD_REQUIRE(a.d_in == nullptr || (reinterpret_cast<uintptr_t>(a.d_in) & 15) == 0,
                "hei_layer_bwd: d_in must be 16-byte aligned");
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    AREAD_CUDA(cudaFuncSetAttribute(hei_tc_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
    AREAD_CUDA(cudaFuncSetAttribute(hei_tc_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
  }
  if (pl.NB == 32) {
    AREAD_LAUNCH(hei_tc_bwd_kernel<32>, pl.grid, kThreads, BwdSmem::kTotal, stream, p);
  } else {
    AREAD_LAUNCH(hei_tc_bwd_kernel<64>, pl.grid, kThreads, BwdSmem::kTotal, stream, p);
  }
  AREAD_LAUNCH(hei_tc_wgrad_reduce_kernel, ceil_div(n_w, 256), 256, 0, stream, p.partial_w, pl.nb_ctas, n_w, a.d_w);
  if (src_bn) {
    aread_bn_act_bwd_args f = {};
    f.m = a.m;
    f.width = src_width;
    f.bn_skip = a.bn_skip;
    f.d_gamma = a.src_d_gamma;
    f.d_beta = a.src_d_beta;
    f.d_bias = a.src_d_bias;
    AREAD_LAUNCH(bn_bwd_finalize_kernel, ceil_div(f.width, 32), kBnThreads, 0, stream, f, p.partial_s, pl.nb_ctas,
                 a.src_coef);
  }
  return AREAD_OK;
}

}  // namespace aread
