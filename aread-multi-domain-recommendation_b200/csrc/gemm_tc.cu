// Grouped Linear layers of the MMoE experts / HEI towers on the 5th-generation tensor cores.
//
//   C[:, g*n : (g+1)*n] = A[:, g*a_group_cols : +k] . W_g^T (+ bias_g)      for every ACTIVE group g
//
// bf16 operands, fp32 accumulation in TMEM.  One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 4-stage ring)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M=128 x N=BN x K=16)
//   warps 2..5  epilogue       (tcgen05.ld -> registers -> bias -> global store), double-buffered
//                              accumulators so the epilogue of tile i overlaps the MMAs of tile i+1
// Groups whose bit is clear in `group_mask` (towers pruned by the HEMP mask) produce no tiles at
// all: the work is skipped, not multiplied by zero.
//
// Reference arithmetic: nn.Linear inside MultiLayerPerceptron (model/layer.py:210, 221-229) as used
// for the experts (model/aread.py:93-95, 150) and towers (aread.py:108-110, 307, 319).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensor_map.cuh"

namespace aread {
namespace {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kGemmThreads = 192;
constexpr int kMaxGroups = 64;

struct GemmParams {
  int64_t m;
  int n, k, n_active;
  int a_group_cols;
  int64_t ldc;
  float* c_f32;
  __nv_bfloat16* c_bf16;
  const float* bias;
  int n_m_tiles, n_n_tiles, n_k_blocks;
  int n_seg;  // 1: bf16 operands; 3: split operands, A_hi.B_hi + A_hi.B_lo + A_lo.B_hi
  int tma_store;  // fp32 output in whole 32-column chunks: the epilogue stages them in smem and stores with TMA
  int64_t total_tiles;
  // fused BatchNorm bookkeeping (aread_expert_gemm); `width` = groups * n
  int width;
  float* partial;            // [n_m_tiles][2][width]
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  const __nv_bfloat16* z;    // BN_BWD: pre-activation of the layer below, [m, ldz]
  int64_t ldz;
  uint32_t threshold;        // dropout of the layer below
  float keep_scale;
  uint32_t salt;
  uint64_t seed;
  const uint64_t* seed_ptr;
  unsigned char group_ids[kMaxGroups];
};

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreOffset = kStages * kStageBytes;           // epilogue staging: 4 warps x 2 x [32 rows][128 B]
  static constexpr int kStoreBytes = 4 * 2 * 32 * 128;
  static constexpr int kPartOffset = kStoreOffset + kStoreBytes;        // column partials of the 4 epilogue warps
  static constexpr int kPartBytes = 2 * 4 * 2 * BN * 4;                 // [2 buffers][4 warps][2][BN] fp32
  static constexpr int kBarrierOffset = kPartOffset + kPartBytes;
  static constexpr int kTotal = kBarrierOffset + 256 + 1024;  // barriers + slack for 1024-byte alignment
};

struct TileCoord {
  int m_t, g, n_t;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int64_t tile) {
  TileCoord c;
  c.n_t = static_cast<int>(tile % p.n_n_tiles);
  const int64_t r = tile / p.n_n_tiles;
  c.g = p.group_ids[r % p.n_active];
  c.m_t = static_cast<int>(r / p.n_active);
  return c;
}

// counter-based dropout stream shared with bn_act.cu / hei.cu (bn_common.cuh): keep(element) is a pure function of
// (seed, salt, element index); one 32-bit hash serves elements 2q and 2q + 1 (16 bits each)
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// pairs at and beyond 2^32 (tensors of 2^33 elements and more): kept out of line so that the common path is one mix32
__device__ __noinline__ uint32_t dropout_inner_far(uint32_t hi, uint32_t ks) { return mix32(hi ^ ks); }
// `seed`, `salt` and a pair below 2^32 make the inner hash a per-launch constant that the compiler hoists out of the
// element loops: one mix32 per pair instead of two, the same bits as before.
__device__ __forceinline__ uint32_t dropout_pair_hash(uint64_t seed, uint32_t salt, uint64_t pair) {
  const uint32_t hi = static_cast<uint32_t>(pair >> 32);
  const uint32_t ks = salt ^ static_cast<uint32_t>(seed);
  uint32_t inner = mix32(ks);                          // loop-invariant
  if (hi != 0u) inner = dropout_inner_far(hi, ks);
  return mix32(static_cast<uint32_t>(pair) ^ inner ^ static_cast<uint32_t>(seed >> 32));
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t salt, uint64_t idx, uint32_t threshold) {
  const uint32_t h = dropout_pair_hash(seed, salt, idx >> 1);
  return ((idx & 1) ? (h >> 16) : (h & 0xffffu)) >= threshold;
}

// Column sums over the 32 rows a warp holds (thread = row, v[j] = column j): a butterfly that halves the number of
// live values at every step, 31 shuffles in all.  Lane L returns the total of column L; the order of the additions
// is fixed.  `v` is destroyed.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = upper ? v[j] : v[j + off];
      const float keep = upper ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kEpiPlain = AREAD_EPI_PLAIN, kEpiStats = AREAD_EPI_STATS, kEpiAct = AREAD_EPI_ACT,
              kEpiBnBwd = AREAD_EPI_BN_BWD, kEpiBf16 = AREAD_EPI_BF16;

template <int BN, int EPI, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
grouped_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_lo,
                      const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kBoxBytes = BK * 64 * 2;   // MN-major B: [64 k rows][64 n] boxes
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  constexpr uint32_t kTmemCols = kAccStages * BN;  // 128 or 256: a power of two >= 32

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord c = decode_tile(p, tile);
        const int a_col0 = c.g * p.a_group_cols;
        const int b_row0 = B_MN ? c.g * p.k : c.g * p.n + c.n_t * BN;
        for (int seg = 0; seg < p.n_seg; ++seg) {
          const CUtensorMap* ma = seg < 2 ? &map_a : &map_a_lo;            // hi.hi, hi.lo, lo.hi, lo.lo
          const CUtensorMap* mb = (seg & 1) ? &map_b_lo : &map_b;
          for (int kb = 0; kb < p.n_k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            ptx::tma_load_2d(sa, ma, &full_bar[stage], a_col0 + kb * BK, c.m_t * BM);
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                ptx::tma_load_2d(sb + j * kBoxBytes, mb, &full_bar[stage], c.n_t * BN + j * 64, b_row0 + kb * BK);
            } else {
              ptx::tma_load_2d(sb, mb, &full_bar[stage], kb * BK, b_row0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_ab(BM, BN, /*a_mn=*/false, /*b_mn=*/B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int n_kb = p.n_k_blocks * p.n_seg;
        for (int kb = 0; kb < n_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t da = ptx::umma_desc_k_sw128(sa + kk * UMMA_K * 2, 8 * 128);
            const uint64_t db = B_MN ? ptx::umma_desc_mn_sw128(sb + kk * 2048, kBoxBytes, 1024)
                                     : ptx::umma_desc_k_sw128(sb + kk * UMMA_K * 2, 8 * 128);
            ptx::umma_bf16(d_tmem, da, db, idesc, (kb | kk) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {  // ===== epilogue warps: TMEM lane quadrant = warp % 4 =====
    const int quad = warp % 4;
    int acc = 0;
    int store_buf = 0;
    int part_buf = 0;
    uint32_t acc_phase = 0;
    uint64_t seed = 0;
    if (EPI == kEpiBnBwd) seed = p.seed_ptr != nullptr ? __ldg(p.seed_ptr) : p.seed;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(p, tile);
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int64_t row = static_cast<int64_t>(c.m_t) * BM + quad * 32 + lane;
      const int n0 = c.n_t * BN;  // column inside the group
      if (EPI == kEpiPlain) {
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          float v[32];
          ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cc * 32, v);
          const int col_in_group = n0 + cc * 32;
          if (col_in_group >= p.n) continue;
          const int64_t col = static_cast<int64_t>(c.g) * p.n + col_in_group;
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col_in_group + j < p.n) v[j] += __ldg(p.bias + col + j);
          }
          if (p.tma_store) {
            // stage the warp's [32 rows][32 fp32] chunk in 128B-swizzled shared memory (what the tensor map expects)
            // and let the TMA unit write whole 128-byte rows; rows past m are clipped by the map
            uint8_t* stage_buf = smem + L::kStoreOffset + (quad * 2 + store_buf) * (32 * 128);
            if (lane == 0) ptx::tma_store_wait_read<1>();       // the store that used this buffer two chunks ago
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&map_c, stage_buf, static_cast<int32_t>(col), c.m_t * BM + quad * 32);
              ptx::tma_store_commit();
            }
            store_buf ^= 1;
            continue;
          }
          if (row < p.m) {
            const bool full = col_in_group + 32 <= p.n;
            if (p.c_f32 != nullptr) {
              float* dst = p.c_f32 + row * p.ldc + col;
              if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int j = 0; j < 32 && col_in_group + j < p.n; ++j) dst[j] = v[j];
              }
            } else {
              __nv_bfloat16* dst = p.c_bf16 + row * p.ldc + col;
              if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk;
                  pk.x = pack_bf16(v[j], v[j + 1]);
                  pk.y = pack_bf16(v[j + 2], v[j + 3]);
                  pk.z = pack_bf16(v[j + 4], v[j + 5]);
                  pk.w = pack_bf16(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(dst + j) = pk;
                }
              } else {
                for (int j = 0; j < 32 && col_in_group + j < p.n; ++j) dst[j] = __float2bfloat16_rn(v[j]);
              }
            }
          }
        }
      } else {
        // ---- bf16 output in [32 rows][64 columns] chunks (128-byte rows, TMA store) with the BatchNorm bookkeeping.
        // n is a multiple of 64 here, so every chunk is whole.
        float* part = reinterpret_cast<float*>(smem + L::kPartOffset) + (part_buf * 4 + quad) * (2 * BN);
#pragma unroll 1
        for (int cc = 0; cc < BN / 64; ++cc) {
          const int col_in_group = n0 + cc * 64;
          if (col_in_group >= p.n) {     // (BN = 128 over a 64-wide group never happens: BN is chosen from n)
            continue;
          }
          const int64_t col = static_cast<int64_t>(c.g) * p.n + col_in_group;
          uint32_t packed[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
            ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cc * 64 + half * 32, v);
            const int64_t hcol = col + half * 32;
            if (EPI == kEpiBf16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) packed[half * 16 + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
            } else if (EPI == kEpiAct) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(fmaf(v[j], __ldg(p.scale + hcol + j), __ldg(p.shift + hcol + j)), 0.f);
#pragma unroll
              for (int j = 0; j < 16; ++j) packed[half * 16 + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
            } else if (EPI == kEpiStats) {
#pragma unroll
              for (int j = 0; j < 16; ++j) packed[half * 16 + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
              float sq[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
              const float s1 = warp_column_sums(v, lane);
              const float s2 = warp_column_sums(sq, lane);
              part[cc * 64 + half * 32 + lane] = s1;
              part[BN + cc * 64 + half * 32 + lane] = s2;
            } else {  // kEpiBnBwd
              float xh[32];
              uint4 zr[4];
              if (row < p.m) {
                const uint4* zp = reinterpret_cast<const uint4*>(p.z + row * p.ldz + hcol);
#pragma unroll
                for (int j = 0; j < 4; ++j) zr[j] = __ldg(zp + j);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) zr[j] = make_uint4(0, 0, 0, 0);
              }
              const uint32_t* zw = reinterpret_cast<const uint32_t*>(zr);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint32_t w = zw[j >> 1];
                const float zf = __uint_as_float((j & 1) ? (w & 0xffff0000u) : (w << 16));
                const float y = fmaf(zf, __ldg(p.scale + hcol + j), __ldg(p.shift + hcol + j));
                const bool keep = p.threshold == 0u ||
                                  dropout_keep(seed, p.salt, static_cast<uint64_t>(row) * p.width + hcol + j, p.threshold);
                const float dy = (y > 0.f && keep && row < p.m) ? v[j] * p.keep_scale : 0.f;
                v[j] = dy;
                xh[j] = dy * (zf - __ldg(p.mean + hcol + j)) * __ldg(p.rstd + hcol + j);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) packed[half * 16 + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
              const float s1 = warp_column_sums(v, lane);
              const float s2 = warp_column_sums(xh, lane);
              part[cc * 64 + half * 32 + lane] = s1;
              part[BN + cc * 64 + half * 32 + lane] = s2;
            }
          }
          uint8_t* stage_buf = smem + L::kStoreOffset + (quad * 2 + store_buf) * (32 * 128);
          if (lane == 0) ptx::tma_store_wait_read<1>();       // the store that used this buffer two chunks ago
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stage_buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&map_c, stage_buf, static_cast<int32_t>(col), c.m_t * BM + quad * 32);
            ptx::tma_store_commit();
          }
          store_buf ^= 1;
        }
        if (EPI == kEpiStats || EPI == kEpiBnBwd) {
          // the four warps' partials -> one partial per 128-row tile, added in warp order
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int t = quad * 32 + lane;
          const float* pb = reinterpret_cast<const float*>(smem + L::kPartOffset) + part_buf * 4 * (2 * BN);
          for (int o = t; o < 2 * BN; o += 128) {
            const int which = o / BN, cj = o % BN;
            if (n0 + cj < p.n) {
              float sum = pb[o];
#pragma unroll
              for (int q = 1; q < 4; ++q) sum += pb[q * (2 * BN) + o];
              p.partial[(static_cast<int64_t>(c.m_t) * 2 + which) * p.width + static_cast<int64_t>(c.g) * p.n + n0 + cj] = sum;
            }
          }
          part_buf ^= 1;
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if ((EPI != kEpiPlain || p.tma_store) && lane == 0) ptx::tma_store_wait_read<0>();   // smem must outlive the last stores
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ----------------------------------------------------------------------------------------------
// Weight gradient: dW_g[n, k] = sum_m dZ[m, g*n_out + n] * A[m, g*a_group_cols + k].
// The reduction runs over the samples, so both operands are consumed in their natural row-major
// layout as MN-major UMMA operands (no transposed copies).  The sample range is split over
// `n_split` CTAs per output tile; fp32 partials are summed in split order by wgrad_reduce_kernel
// (deterministic).  Output tile: 128 output features x BN input features.
// ----------------------------------------------------------------------------------------------
struct WgradParams {
  int64_t m;
  int n, k, n_active, n_split;
  int a_group_cols;
  float* partial;            // [n_split][groups_active][n][k]
  int n_n_tiles, n_k_tiles;  // tiles over output features / input features
  int kb_per_split, n_kb;    // 64-sample blocks
  int n_seg;                 // 1: bf16; 3: dZ_hi.A_hi + dZ_hi.A_lo + dZ_lo.A_hi
  int64_t total_tiles;
  unsigned char group_ids[kMaxGroups];
};

struct WgradCoord {
  int split, gi, g, n_t, k_t;
};

__device__ __forceinline__ WgradCoord decode_wgrad(const WgradParams& p, int64_t tile) {
  WgradCoord c;
  c.k_t = static_cast<int>(tile % p.n_k_tiles);
  int64_t r = tile / p.n_k_tiles;
  c.n_t = static_cast<int>(r % p.n_n_tiles);
  r /= p.n_n_tiles;
  c.gi = static_cast<int>(r % p.n_active);
  c.g = p.group_ids[c.gi];
  c.split = static_cast<int>(r / p.n_active);
  return c;
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
grouped_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_dz_lo, const __grid_constant__ CUtensorMap map_a_lo,
                     const WgradParams p) {
  constexpr int kBoxBytes = BK * 64 * 2;                 // [64 samples][64 features] bf16 = 8 KB
  constexpr int kABytes = 2 * kBoxBytes;                 // 128 output features
  constexpr int kBBytes = (BN / 64) * kBoxBytes;         // BN input features
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  constexpr uint32_t kTmemCols = kAccStages * BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_dz);
    ptx::prefetch_tensormap(&map_a);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const WgradCoord c = decode_wgrad(p, tile);
        const int kb0 = c.split * p.kb_per_split;
        const int kb1 = min(p.n_kb, kb0 + p.kb_per_split);
        const int dz_col0 = c.g * p.n + c.n_t * BM;
        const int a_col0 = c.g * p.a_group_cols + c.k_t * BN;
        for (int seg = 0; seg < p.n_seg; ++seg) {
          const CUtensorMap* mz = seg < 2 ? &map_dz : &map_dz_lo;
          const CUtensorMap* ma = seg == 1 ? &map_a_lo : &map_a;
          for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kStageBytes;
            uint8_t* sb = sa + kABytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            ptx::tma_load_2d(sa, mz, &full_bar[stage], dz_col0, kb * BK);
            ptx::tma_load_2d(sa + kBoxBytes, mz, &full_bar[stage], dz_col0 + 64, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * kBoxBytes, ma, &full_bar[stage], a_col0 + j * 64, kb * BK);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, /*mn_major=*/true);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const WgradCoord c = decode_wgrad(p, tile);
        const int kb0 = c.split * p.kb_per_split;
        const int kb1 = min(p.n_kb, kb0 + p.kb_per_split);
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int n_iter = (kb1 - kb0) * p.n_seg;
        for (int it = 0; it < n_iter; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {  // 16 samples = two 8-row swizzle atoms
            const uint64_t da = ptx::umma_desc_mn_sw128(sa + kk * 2048, kBoxBytes, 1024);
            const uint64_t db = ptx::umma_desc_mn_sw128(sb + kk * 2048, kBoxBytes, 1024);
            ptx::umma_bf16(d_tmem, da, db, idesc, (it > 0) || (kk != 0));
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&acc_full[acc]);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {  // ===== epilogue: fp32 partial tile =====
    const int quad = warp % 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const WgradCoord c = decode_wgrad(p, tile);
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int row = c.n_t * BM + quad * 32 + lane;  // output feature inside the group
      float* base = p.partial + ((static_cast<int64_t>(c.split) * p.n_active + c.gi) * p.n + row) * p.k;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        float v[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cc * 32, v);
        const int col = c.k_t * BN + cc * 32;
        if (row < p.n && col < p.k) {
          float* dst = base + col;
          if (col + 32 <= p.k && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32 && col + j < p.k; ++j) dst[j] = v[j];
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// dw[g][n][k] = sum_s partial[s][gi][n][k] in split order; inactive groups are left untouched
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           int n_split, int n_active, int64_t group_elems,
                                                           const WgradParams p) {
  const int64_t total = static_cast<int64_t>(n_active) * group_elems;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int gi = static_cast<int>(i / group_elems);
    const int64_t e = i - gi * group_elems;
    float acc = partial[i];
    for (int s = 1; s < n_split; ++s) acc += partial[static_cast<int64_t>(s) * total + i];
    dw[static_cast<int64_t>(p.group_ids[gi]) * group_elems + e] = acc;
  }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ----------------------------------------------------------------------------------------------
inline unsigned gemm_grid(int64_t tiles) {
  const int n = sm_count();
  return static_cast<unsigned>(tiles < n ? tiles : n);     // persistent: one CTA per SM
}

template <int BN, int EPI, bool B_MN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma_lo, const CUtensorMap& mb_lo,
                const CUtensorMap& mc, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static uint64_t configured = 0;
  if (first_use_on_device(&configured))
    AREAD_CUDA(cudaFuncSetAttribute(grouped_linear_kernel<BN, EPI, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kTotal));
  AREAD_LAUNCH((grouped_linear_kernel<BN, EPI, B_MN>), gemm_grid(p.total_tiles), kGemmThreads, L::kTotal, stream, ma, mb,
               ma_lo, mb_lo, mc, p);
  return AREAD_OK;
}


template <int BN>
int launch_wgrad(const CUtensorMap& mdz, const CUtensorMap& ma, const CUtensorMap& mdz_lo, const CUtensorMap& ma_lo,
                 const WgradParams& p, cudaStream_t stream) {
  constexpr int kSmem = kStages * (2 + BN / 64) * (BK * 64 * 2) + 256 + 1024;
  static uint64_t configured = 0;
  if (first_use_on_device(&configured))
    AREAD_CUDA(cudaFuncSetAttribute(grouped_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  AREAD_LAUNCH((grouped_wgrad_kernel<BN>), gemm_grid(p.total_tiles), kGemmThreads, kSmem, stream, mdz, ma, mdz_lo, ma_lo,
               p);
  return AREAD_OK;
}

int wgrad_plan(const aread_grouped_wgrad_args& a, WgradParams* p, int* bn_out) {
  p->m = a.m;
  p->n = a.n;
  p->k = a.k;
  p->a_group_cols = a.a_group_cols;
  p->n_active = 0;
  for (int g = 0; g < a.groups; ++g)
    if (a.group_mask & (uint64_t{1} << g)) p->group_ids[p->n_active++] = static_cast<unsigned char>(g);
  const int bn = a.k > 64 ? 128 : 64;
  *bn_out = bn;
  p->n_n_tiles = ceil_div(a.n, BM);
  p->n_k_tiles = ceil_div(a.k, bn);
  p->n_kb = ceil_div(a.m, BK);
  const int64_t out_tiles = static_cast<int64_t>(p->n_active > 0 ? p->n_active : 1) * p->n_n_tiles * p->n_k_tiles;
  int split = static_cast<int>((2 * kNumSMs + out_tiles - 1) / out_tiles);  // about two waves of CTAs
  if (split > p->n_kb) split = p->n_kb;
  if (split < 1) split = 1;
  p->kb_per_split = ceil_div(p->n_kb, split);
  p->n_split = ceil_div(p->n_kb, p->kb_per_split);
  p->total_tiles = out_tiles * p->n_split;
  return AREAD_OK;
}

}  // namespace
}  // namespace aread

extern "C" int aread_grouped_linear_bf16(const aread_grouped_linear_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "grouped_linear: null args");
  const aread_grouped_linear_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.n > 0 && a.k > 0, "grouped_linear: bad shape m=%lld n=%d k=%d", (long long)a.m, a.n, a.k);
  AREAD_REQUIRE(a.groups > 0 && a.groups <= kMaxGroups, "grouped_linear: groups %d not in [1, %d]", a.groups,
                kMaxGroups);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.a && a.b, "grouped_linear: null operand");
  AREAD_REQUIRE((a.c_f32 != nullptr) != (a.c_bf16 != nullptr), "grouped_linear: exactly one of c_f32 / c_bf16");
  AREAD_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "grouped_linear: lda/ldb must be multiples of 8 bf16 (16 bytes)");
  AREAD_REQUIRE(reinterpret_cast<uintptr_t>(a.a) % 16 == 0 && reinterpret_cast<uintptr_t>(a.b) % 16 == 0,
                "grouped_linear: operands must be 16-byte aligned");
  AREAD_REQUIRE(a.a_group_cols == 0 || a.a_group_cols >= a.k, "grouped_linear: a_group_cols %d < k %d",
                a.a_group_cols, a.k);

  GemmParams p{};
  p.m = a.m;
  p.n = a.n;
  p.k = a.k;
  p.a_group_cols = a.a_group_cols;
  p.ldc = a.ldc;
  p.c_f32 = a.c_f32;
  p.c_bf16 = reinterpret_cast<__nv_bfloat16*>(a.c_bf16);
  p.bias = a.bias;
  p.n_active = 0;
  for (int g = 0; g < a.groups; ++g)
    if (a.group_mask & (uint64_t{1} << g)) p.group_ids[p.n_active++] = static_cast<unsigned char>(g);
  if (p.n_active == 0) return AREAD_OK;
  const int bn = a.n > 64 ? 128 : 64;
  p.n_m_tiles = ceil_div(a.m, BM);
  p.n_n_tiles = ceil_div(a.n, bn);
  p.n_k_blocks = ceil_div(a.k, BK);
  p.total_tiles = static_cast<int64_t>(p.n_m_tiles) * p.n_active * p.n_n_tiles;

  AREAD_REQUIRE((a.a_lo != nullptr) == (a.b_lo != nullptr), "grouped_linear: a_lo and b_lo go together");
  p.n_seg = a.a_lo != nullptr ? (a.lo_lo ? 4 : 3) : 1;
  CUtensorMap ma, mb, ma_lo, mb_lo;
  const int64_t a_cols = a.a_group_cols == 0 ? a.k : static_cast<int64_t>(a.a_group_cols) * (a.groups - 1) + a.k;
  if (int rc = make_map(&ma, a.a, a.m, a_cols, a.lda, BM)) return rc;
  if (int rc = make_map(&mb, a.b, static_cast<int64_t>(a.groups) * a.n, a.k, a.ldb, bn)) return rc;
  ma_lo = ma;
  mb_lo = mb;
  if (p.n_seg >= 3) {
    AREAD_REQUIRE(reinterpret_cast<uintptr_t>(a.a_lo) % 16 == 0 && reinterpret_cast<uintptr_t>(a.b_lo) % 16 == 0,
                  "grouped_linear: lo operands must be 16-byte aligned");
    if (int rc = make_map(&ma_lo, a.a_lo, a.m, a_cols, a.lda, BM)) return rc;
    if (int rc = make_map(&mb_lo, a.b_lo, static_cast<int64_t>(a.groups) * a.n, a.k, a.ldb, bn)) return rc;
  }
  CUtensorMap mc = ma;
  p.tma_store = (a.c_f32 != nullptr && a.n % 32 == 0 && a.ldc % 4 == 0 && reinterpret_cast<uintptr_t>(a.c_f32) % 16 == 0)
                    ? 1 : 0;
  if (p.tma_store)
    if (int rc = make_store_map(&mc, a.c_f32, a.m, static_cast<int64_t>(a.groups) * a.n, a.ldc)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  return bn == 128 ? launch_gemm<128, kEpiPlain, false>(ma, mb, ma_lo, mb_lo, mc, p, stream)
                   : launch_gemm<64, kEpiPlain, false>(ma, mb, ma_lo, mb_lo, mc, p, stream);
}

// ----------------------------------------------------------------------------------------------
// aread_expert_gemm: the same pipeline with the BatchNorm bookkeeping in the epilogue
// ----------------------------------------------------------------------------------------------
namespace aread {
namespace {

// bf16 row-major [rows, cols] output, box = [32 rows, 64 cols] (one 128-byte swizzle row per matrix row)
int make_store_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 store) failed with CUresult %d", (int)r);
  return AREAD_OK;
}

inline uint32_t dropout_threshold_of(float p) {   // 16-bit scale (bn_common.cuh dropout_threshold)
  if (p <= 0.f) return 0u;
  const double t = static_cast<double>(p) * 65536.0;
  return t >= 65535.0 ? 0xffffu : static_cast<uint32_t>(t + 0.5);
}

template <int EPI, bool B_MN>
int launch_expert(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const GemmParams& p,
                  cudaStream_t stream) {
  return bn == 128 ? launch_gemm<128, EPI, B_MN>(ma, mb, ma, mb, mc, p, stream)
                   : launch_gemm<64, EPI, B_MN>(ma, mb, ma, mb, mc, p, stream);
}

__global__ void __launch_bounds__(1024) expert_bn_finalize_kernel(const aread_expert_bn_finalize_args a) {
  // 32 columns x 32 partial groups per CTA; partials are added per thread in tile order, then in thread order
  __shared__ float s_c[2][32][33];
  const int cx = threadIdx.x % 32, py = threadIdx.x / 32;
  const int col = blockIdx.x * 32 + cx;
  float s = 0.f, q = 0.f;
  if (a.training && !a.bn_skip && col < a.width) {
    for (int i = py; i < a.n_partial; i += 32) {
      s += a.partial[(static_cast<int64_t>(i) * 2 + 0) * a.width + col];
      q += a.partial[(static_cast<int64_t>(i) * 2 + 1) * a.width + col];
    }
  }
  s_c[0][py][cx] = s;
  s_c[1][py][cx] = q;
  __syncthreads();
  if (py != 0 || col >= a.width) return;
  s = q = 0.f;
  for (int y = 0; y < 32; ++y) { s += s_c[0][y][cx]; q += s_c[1][y][cx]; }
  const float bias = a.bias != nullptr ? a.bias[col] : 0.f;
  float mean, rstd, scale, shift;
  if (a.bn_skip) {                       // identity instead of BatchNorm: h = relu(acc + bias)
    mean = 0.f; rstd = 1.f; scale = 1.f; shift = bias;
  } else if (a.training) {
    const float inv_m = 1.f / static_cast<float>(a.m);
    mean = s * inv_m;                    // of the bias-free accumulator
    const float var = fmaxf(q * inv_m - mean * mean, 0.f);
    rstd = 1.f / sqrtf(var + a.eps);
    scale = a.gamma[col] * rstd;
    shift = a.beta[col] - mean * scale;
    const float unbiased = a.m > 1 ? var * (static_cast<float>(a.m) / static_cast<float>(a.m - 1)) : var;
    a.running_mean[col] = (1.f - a.momentum) * a.running_mean[col] + a.momentum * (mean + bias);
    a.running_var[col] = (1.f - a.momentum) * a.running_var[col] + a.momentum * unbiased;
  } else {
    rstd = 1.f / sqrtf(a.running_var[col] + a.eps);
    mean = a.running_mean[col] - bias;
    scale = a.gamma[col] * rstd;
    shift = a.beta[col] - mean * scale;
  }
  a.mean[col] = mean;
  a.rstd[col] = rstd;
  a.scale[col] = scale;
  a.shift[col] = shift;
}

__global__ void __launch_bounds__(1024) expert_bn_bwd_finalize_kernel(const aread_expert_bn_bwd_finalize_args a) {
  __shared__ float s_c[2][32][33];
  const int cx = threadIdx.x % 32, py = threadIdx.x / 32;
  const int col = blockIdx.x * 32 + cx;
  float s1 = 0.f, s2 = 0.f;
  if (col < a.width) {
    for (int i = py; i < a.n_partial; i += 32) {
      s1 += a.partial[(static_cast<int64_t>(i) * 2 + 0) * a.width + col];
      s2 += a.partial[(static_cast<int64_t>(i) * 2 + 1) * a.width + col];
    }
  }
  s_c[0][py][cx] = s1;
  s_c[1][py][cx] = s2;
  __syncthreads();
  if (py != 0 || col >= a.width) return;
  s1 = s2 = 0.f;
  for (int y = 0; y < 32; ++y) { s1 += s_c[0][y][cx]; s2 += s_c[1][y][cx]; }
  if (a.bn_skip) {
    if (a.d_gamma) a.d_gamma[col] = 0.f;
    if (a.d_beta) a.d_beta[col] = 0.f;
    if (a.d_bias) a.d_bias[col] = s1;
    a.coef[col] = 0.f;
    a.coef[a.width + col] = 0.f;
  } else {
    if (a.d_gamma) a.d_gamma[col] = s2;
    if (a.d_beta) a.d_beta[col] = s1;
    if (a.d_bias) a.d_bias[col] = 0.f;  // BatchNorm removes the column mean: the exact gradient is zero
    const float inv_m = 1.f / static_cast<float>(a.m);
    a.coef[col] = s1 * inv_m;
    a.coef[a.width + col] = s2 * inv_m;
  }
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// BatchNorm + ReLU + Dropout passes over a bf16 pre-activation.  A thread owns 8 consecutive columns (one 16-byte
// access) and walks rows, so the per-column constants live in registers; the CTA owns a contiguous range of rows.
//   MODE 0  forward:   out = bf16(dropout(relu(z * scale + shift)))
//   MODE 1  backward:  dz = bf16(scale * (dy - coef0 - xhat * coef1)), dy masked here when it arrives raw
//   MODE 2  backward statistics: per-CTA partials of sum(dy) and sum(dy * xhat) (rows in order: deterministic)
constexpr int kBn16Threads = 256;

template <int MODE>
__global__ void __launch_bounds__(kBn16Threads, 2) bn16_kernel(const aread_bn16_args a, uint32_t threshold, float keep_scale,
                                                               int tpr, int64_t rows_per_cta, float* __restrict__ partial) {
  __shared__ float s_red[MODE == 2 ? 2 * kBn16Threads * 8 : 1];
  const uint64_t seed = a.seed_ptr != nullptr ? __ldg(a.seed_ptr) : a.seed;
  const int cg = a.width / 8;
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr, ty_n = kBn16Threads / tpr;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(a.m, r0 + rows_per_cta);
  const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(a.z);
  const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(a.dy);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.out);
  // the ReLU / dropout pattern of a row's 8 columns is one byte: written by the forward when the caller keeps it,
  // read back by the backward passes instead of hashing again
  const bool use_bits = MODE != 0 && a.dy_is_raw != 0 && a.pass_bits != nullptr;
  const bool mask_here = MODE == 0 || (a.dy_is_raw != 0 && !use_bits);
  for (int g0 = 0; g0 < cg; g0 += tpr) {
    const int g = g0 + tx;
    const bool on = g < cg && ty < ty_n;
    const int col = g * 8;
    float sc[8], sh[8], mu[8], rs[8], c0[8], c1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = on ? __ldg(a.scale + col + j) : 0.f;
      sh[j] = (on && mask_here) ? __ldg(a.shift + col + j) : 0.f;
      if (MODE != 0) {
        mu[j] = (on && !a.bn_skip) ? __ldg(a.mean + col + j) : 0.f;
        rs[j] = (on && !a.bn_skip) ? __ldg(a.rstd + col + j) : 0.f;
      }
      if (MODE == 1) {
        c0[j] = (on && !a.bn_skip) ? __ldg(a.coef + col + j) : 0.f;
        c1[j] = (on && !a.bn_skip) ? __ldg(a.coef + a.width + col + j) : 0.f;
      }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (on) {
      constexpr int U = 4;                      // rows in flight per thread: the loads of a batch are issued together
      for (int64_t rb = r0 + ty; rb < r1; rb += static_cast<int64_t>(ty_n) * U) {
        uint4 zq[U], dq[U];
        uint32_t bq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = rb + static_cast<int64_t>(u) * ty_n;
          zq[u] = dq[u] = make_uint4(0, 0, 0, 0);
          bq[u] = 0u;
          if (r < r1) {
            zq[u] = __ldg(reinterpret_cast<const uint4*>(zb + r * a.ldz + col));
            if (MODE != 0) dq[u] = __ldg(reinterpret_cast<const uint4*>(db + r * a.ldd + col));
            if (use_bits) bq[u] = a.pass_bits[r * cg + g];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = rb + static_cast<int64_t>(u) * ty_n;
          if (r >= r1) break;
          const uint32_t zw[4] = {zq[u].x, zq[u].y, zq[u].z, zq[u].w};
          const uint32_t dw[4] = {dq[u].x, dq[u].y, dq[u].z, dq[u].w};
          const uint32_t bits = bq[u];
          uint32_t bits_out = 0u;
          uint32_t hsh[4] = {0u, 0u, 0u, 0u};
          if (mask_here && threshold != 0u) {
            const uint64_t pair0 = (static_cast<uint64_t>(r) * a.width + col) >> 1;     // col % 8 == 0, width % 8 == 0
#pragma unroll
            for (int q = 0; q < 4; ++q) hsh[q] = dropout_pair_hash(seed, a.salt, pair0 + q);
          }
          float out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = (j & 1) ? bf16_hi(zw[j >> 1]) : bf16_lo(zw[j >> 1]);
            bool pass = true;
            float y = 0.f;
            if (mask_here) {
              y = fmaf(z, sc[j], sh[j]);
              pass = y > 0.f && (threshold == 0u || ((j & 1) ? (hsh[j >> 1] >> 16) : (hsh[j >> 1] & 0xffffu)) >= threshold);
              bits_out |= pass ? (1u << j) : 0u;
            } else if (use_bits) {
              pass = (bits >> j) & 1u;
            }
            if (MODE == 0) {
              out[j] = pass ? y * keep_scale : 0.f;
            } else {
              float dy = (j & 1) ? bf16_hi(dw[j >> 1]) : bf16_lo(dw[j >> 1]);
              if (a.dy_is_raw) dy = pass ? dy * a.keep_scale_bwd : 0.f;
              const float xhat = (z - mu[j]) * rs[j];
              if (MODE == 1) {
                out[j] = a.bn_skip ? dy : sc[j] * (dy - c0[j] - xhat * c1[j]);
              } else {
                s1[j] += dy;
                s2[j] = fmaf(dy, xhat, s2[j]);
              }
            }
          }
          if (MODE == 0 && a.pass_bits != nullptr) a.pass_bits[r * cg + g] = static_cast<uint8_t>(bits_out);
          if (MODE != 2) {
            uint4 pk;
            pk.x = pack_bf16(out[0], out[1]);
            pk.y = pack_bf16(out[2], out[3]);
            pk.z = pack_bf16(out[4], out[5]);
            pk.w = pack_bf16(out[6], out[7]);
            *reinterpret_cast<uint4*>(ob + r * a.ldo + col) = pk;
          }
        }
      }
    }
    if (MODE == 2) {   // the row lanes of the CTA, added in lane order
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_red[(threadIdx.x * 2 + 0) * 8 + j] = s1[j];
        s_red[(threadIdx.x * 2 + 1) * 8 + j] = s2[j];
      }
      __syncthreads();
      if (ty == 0 && g < cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t1 = 0.f, t2 = 0.f;
          for (int y = 0; y < ty_n; ++y) {
            t1 += s_red[((y * tpr + tx) * 2 + 0) * 8 + j];
            t2 += s_red[((y * tpr + tx) * 2 + 1) * 8 + j];
          }
          partial[(static_cast<int64_t>(blockIdx.x) * 2 + 0) * a.width + col + j] = t1;
          partial[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * a.width + col + j] = t2;
        }
      }
    }
  }
}

struct Bn16Grid {
  int tpr, n_cta;
  int64_t rows_per_cta;
};
inline Bn16Grid bn16_grid(int64_t m, int width) {
  Bn16Grid g;
  const int cg = width / 8;
  g.tpr = 1;
  while (g.tpr < cg && g.tpr < kBn16Threads) g.tpr <<= 1;
  const int ty_n = kBn16Threads / g.tpr;
  int64_t n = (m + static_cast<int64_t>(ty_n) * 16 - 1) / (static_cast<int64_t>(ty_n) * 16);   // >= 16 rows per lane
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 4;
  if (n > cap) n = cap;
  if (n < 1) n = 1;
  g.rows_per_cta = (m + n - 1) / n;
  g.n_cta = static_cast<int>((m + g.rows_per_cta - 1) / g.rows_per_cta);
  return g;
}

}  // namespace
}  // namespace aread

extern "C" int32_t aread_expert_gemm_partials(int64_t m) { return aread::ceil_div(m > 0 ? m : 1, aread::BM); }

extern "C" int aread_expert_gemm(const aread_expert_gemm_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "expert_gemm: null args");
  const aread_expert_gemm_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.n > 0 && a.k > 0, "expert_gemm: bad shape m=%lld n=%d k=%d", (long long)a.m, a.n, a.k);
  AREAD_REQUIRE(a.groups > 0 && a.groups <= kMaxGroups, "expert_gemm: groups %d not in [1, %d]", a.groups, kMaxGroups);
  AREAD_REQUIRE(a.epilogue >= AREAD_EPI_PLAIN && a.epilogue <= AREAD_EPI_BF16, "expert_gemm: unknown epilogue %d",
                a.epilogue);
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.a && a.b, "expert_gemm: null operand");
  AREAD_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "expert_gemm: lda/ldb must be multiples of 8 bf16 (16 bytes)");
  AREAD_REQUIRE(reinterpret_cast<uintptr_t>(a.a) % 16 == 0 && reinterpret_cast<uintptr_t>(a.b) % 16 == 0,
                "expert_gemm: operands must be 16-byte aligned");
  AREAD_REQUIRE(a.a_group_cols == 0 || a.a_group_cols >= a.k, "expert_gemm: a_group_cols %d < k %d", a.a_group_cols, a.k);
  AREAD_REQUIRE(!a.b_is_k_by_n || a.groups == 1 || a.k % BK == 0,
                "expert_gemm: a grouped k-by-n weight needs k %% 64 == 0 (k = %d)", a.k);
  const bool plain = a.epilogue == AREAD_EPI_PLAIN;
  if (plain) {
    AREAD_REQUIRE((a.c_f32 != nullptr) != (a.c_bf16 != nullptr), "expert_gemm: exactly one of c_f32 / c_bf16");
  } else {
    AREAD_REQUIRE(a.c_bf16 != nullptr && a.c_f32 == nullptr, "expert_gemm: this epilogue writes bf16");
    AREAD_REQUIRE(a.n % 64 == 0 && a.ldc % 8 == 0 && reinterpret_cast<uintptr_t>(a.c_bf16) % 16 == 0,
                  "expert_gemm: fused epilogues need n %% 64 == 0 and a 16-byte aligned output");
  }
  if (a.epilogue == AREAD_EPI_STATS || a.epilogue == AREAD_EPI_BN_BWD)
    AREAD_REQUIRE(a.partial != nullptr, "expert_gemm: null partial");
  if (a.epilogue == AREAD_EPI_ACT || a.epilogue == AREAD_EPI_BN_BWD)
    AREAD_REQUIRE(a.scale && a.shift, "expert_gemm: null scale / shift");
  if (a.epilogue == AREAD_EPI_BN_BWD) {
    AREAD_REQUIRE(a.mean && a.rstd && a.z, "expert_gemm: null BN_BWD input");
    AREAD_REQUIRE(a.ldz % 8 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0, "expert_gemm: z must be 16-byte aligned");
    AREAD_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, "expert_gemm: dropout %f not in [0, 1)", a.dropout_p);
  }

  GemmParams p{};
  p.m = a.m;
  p.n = a.n;
  p.k = a.k;
  p.a_group_cols = a.a_group_cols;
  p.ldc = a.ldc;
  p.c_f32 = a.c_f32;
  p.c_bf16 = reinterpret_cast<__nv_bfloat16*>(a.c_bf16);
  p.bias = plain ? a.bias : nullptr;
  p.n_active = a.groups;
  for (int g = 0; g < a.groups; ++g) p.group_ids[g] = static_cast<unsigned char>(g);
  const int bn = a.n > 64 ? 128 : 64;
  p.n_m_tiles = ceil_div(a.m, BM);
  p.n_n_tiles = ceil_div(a.n, bn);
  p.n_k_blocks = ceil_div(a.k, BK);
  p.total_tiles = static_cast<int64_t>(p.n_m_tiles) * p.n_active * p.n_n_tiles;
  p.n_seg = 1;
  p.width = a.groups * a.n;
  p.partial = a.partial;
  p.scale = a.scale;
  p.shift = a.shift;
  p.mean = a.mean;
  p.rstd = a.rstd;
  p.z = reinterpret_cast<const __nv_bfloat16*>(a.z);
  p.ldz = a.ldz;
  p.threshold = dropout_threshold_of(a.dropout_p);
  p.keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  p.salt = a.salt;
  p.seed = a.seed;
  p.seed_ptr = a.seed_ptr;

  CUtensorMap ma, mb, mc;
  const int64_t a_cols = a.a_group_cols == 0 ? a.k : static_cast<int64_t>(a.a_group_cols) * (a.groups - 1) + a.k;
  if (int rc = make_map(&ma, a.a, a.m, a_cols, a.lda, BM)) return rc;
  if (a.b_is_k_by_n) {
    if (int rc = make_map(&mb, a.b, static_cast<int64_t>(a.groups) * a.k, a.n, a.ldb, BK, 64)) return rc;
  } else {
    if (int rc = make_map(&mb, a.b, static_cast<int64_t>(a.groups) * a.n, a.k, a.ldb, bn)) return rc;
  }
  mc = ma;
  if (plain) {
    p.tma_store = (a.c_f32 != nullptr && a.n % 32 == 0 && a.ldc % 4 == 0 && reinterpret_cast<uintptr_t>(a.c_f32) % 16 == 0)
                      ? 1 : 0;
    if (p.tma_store)
      if (int rc = make_store_map(&mc, a.c_f32, a.m, static_cast<int64_t>(a.groups) * a.n, a.ldc)) return rc;
  } else {
    if (int rc = make_store_map_bf16(&mc, a.c_bf16, a.m, static_cast<int64_t>(a.groups) * a.n, a.ldc)) return rc;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool mn = a.b_is_k_by_n != 0;
  switch (a.epilogue) {
    case AREAD_EPI_PLAIN:
      return mn ? launch_expert<kEpiPlain, true>(bn, ma, mb, mc, p, stream)
                : launch_expert<kEpiPlain, false>(bn, ma, mb, mc, p, stream);
    case AREAD_EPI_STATS:
      AREAD_REQUIRE(!mn, "expert_gemm: STATS runs on the forward weight layout");
      return launch_expert<kEpiStats, false>(bn, ma, mb, mc, p, stream);
    case AREAD_EPI_ACT:
      AREAD_REQUIRE(!mn, "expert_gemm: ACT runs on the forward weight layout");
      return launch_expert<kEpiAct, false>(bn, ma, mb, mc, p, stream);
    case AREAD_EPI_BF16:
      return mn ? launch_expert<kEpiBf16, true>(bn, ma, mb, mc, p, stream)
                : launch_expert<kEpiBf16, false>(bn, ma, mb, mc, p, stream);
    default:
      AREAD_REQUIRE(mn, "expert_gemm: BN_BWD runs on the k-by-n weight layout");
      return launch_expert<kEpiBnBwd, true>(bn, ma, mb, mc, p, stream);
  }
}

extern "C" int aread_expert_bn_finalize(const aread_expert_bn_finalize_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "expert_bn_finalize: null args");
  const aread_expert_bn_finalize_args& a = *args;
  AREAD_REQUIRE(a.width > 0 && a.m > 0, "expert_bn_finalize: bad shape");
  AREAD_REQUIRE(a.mean && a.rstd && a.scale && a.shift, "expert_bn_finalize: null output");
  AREAD_REQUIRE(a.bn_skip || (a.gamma && a.beta && a.running_mean && a.running_var), "expert_bn_finalize: null BN tensor");
  AREAD_REQUIRE(!(a.training && !a.bn_skip) || (a.partial && a.n_partial > 0), "expert_bn_finalize: null partial");
  AREAD_LAUNCH(expert_bn_finalize_kernel, ceil_div(a.width, 32), 1024, 0, static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

extern "C" int aread_expert_bn_bwd_finalize(const aread_expert_bn_bwd_finalize_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "expert_bn_bwd_finalize: null args");
  const aread_expert_bn_bwd_finalize_args& a = *args;
  AREAD_REQUIRE(a.width > 0 && a.m > 0 && a.n_partial > 0 && a.partial && a.coef, "expert_bn_bwd_finalize: bad arguments");
  AREAD_LAUNCH(expert_bn_bwd_finalize_kernel, ceil_div(a.width, 32), 1024, 0, static_cast<cudaStream_t>(stream_), a);
  return AREAD_OK;
}

extern "C" int32_t aread_bn16_partials(int64_t m, int32_t width) {
  return width > 0 && width % 8 == 0 ? aread::bn16_grid(m > 0 ? m : 1, width).n_cta : 0;
}

namespace aread {
namespace {
int check_bn16(const aread_bn16_args& a, const char* who) {
  AREAD_REQUIRE(a.m >= 0 && a.width > 0 && a.width % 8 == 0, "%s: width %d must be a multiple of 8", who, a.width);
  AREAD_REQUIRE(a.z && a.scale, "%s: null pointer", who);
  AREAD_REQUIRE(a.ldz % 8 == 0 && reinterpret_cast<uintptr_t>(a.z) % 16 == 0, "%s: rows must be 16-byte aligned", who);
  AREAD_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, "%s: dropout %f not in [0, 1)", who, a.dropout_p);
  return AREAD_OK;
}
}  // namespace
}  // namespace aread

extern "C" int aread_bn16(const aread_bn16_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "bn16: null args");
  const aread_bn16_args& a = *args;
  if (int rc = check_bn16(a, "bn16")) return rc;
  if (a.m == 0) return AREAD_OK;
  AREAD_REQUIRE(a.out && a.ldo % 8 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0, "bn16: bad output");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Bn16Grid g = bn16_grid(a.m, a.width);
  const uint32_t threshold = dropout_threshold_of(a.dropout_p);
  const float keep_scale = a.dropout_p > 0.f ? 1.f / (1.f - a.dropout_p) : 1.f;
  if (a.dy != nullptr) {
    AREAD_REQUIRE(a.bn_skip || (a.mean && a.rstd && a.coef), "bn16: null backward input");
    AREAD_REQUIRE(!a.dy_is_raw || a.shift, "bn16: a raw gradient needs scale / shift to rebuild the ReLU mask");
    AREAD_REQUIRE(a.ldd % 8 == 0 && reinterpret_cast<uintptr_t>(a.dy) % 16 == 0, "bn16: dy rows must be 16-byte aligned");
    aread_bn16_args b = a;
    b.keep_scale_bwd = a.dy_is_raw ? keep_scale : 1.f;
    AREAD_LAUNCH(bn16_kernel<1>, g.n_cta, kBn16Threads, 0, stream, b, a.dy_is_raw ? threshold : 0u, 1.f, g.tpr,
                 g.rows_per_cta, nullptr);
  } else {
    AREAD_REQUIRE(a.shift != nullptr, "bn16: null shift");
    AREAD_LAUNCH(bn16_kernel<0>, g.n_cta, kBn16Threads, 0, stream, a, threshold, keep_scale, g.tpr, g.rows_per_cta,
                 nullptr);
  }
  return AREAD_OK;
}

extern "C" int aread_bn16_bwd_stats(const aread_bn16_args* args, float* partial, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr && partial != nullptr, "bn16_bwd_stats: null args");
  const aread_bn16_args& a = *args;
  if (int rc = check_bn16(a, "bn16_bwd_stats")) return rc;
  AREAD_REQUIRE(a.m > 0 && a.dy && a.shift && (a.bn_skip || (a.mean && a.rstd)), "bn16_bwd_stats: null pointer");
  AREAD_REQUIRE(a.ldd % 8 == 0 && reinterpret_cast<uintptr_t>(a.dy) % 16 == 0, "bn16_bwd_stats: dy rows must be 16-byte aligned");
  const Bn16Grid g = bn16_grid(a.m, a.width);
  const uint32_t threshold = a.dy_is_raw ? dropout_threshold_of(a.dropout_p) : 0u;
  const float keep_scale = (a.dy_is_raw && a.dropout_p > 0.f) ? 1.f / (1.f - a.dropout_p) : 1.f;
  aread_bn16_args b = a;
  b.keep_scale_bwd = keep_scale;
  AREAD_LAUNCH(bn16_kernel<2>, g.n_cta, kBn16Threads, 0, static_cast<cudaStream_t>(stream_), b, threshold, 1.f, g.tpr,
               g.rows_per_cta, partial);
  return AREAD_OK;
}

extern "C" size_t aread_grouped_wgrad_workspace_bytes(const aread_grouped_wgrad_args* args) {
  using namespace aread;
  if (args == nullptr || args->m <= 0) return 256;
  WgradParams p{};
  int bn;
  wgrad_plan(*args, &p, &bn);
  return align_up(static_cast<size_t>(p.n_split) * (p.n_active > 0 ? p.n_active : 1) * args->n * args->k * 4, 256);
}

extern "C" int aread_grouped_wgrad_bf16(const aread_grouped_wgrad_args* args, aread_stream_t stream_) {
  using namespace aread;
  AREAD_REQUIRE(args != nullptr, "grouped_wgrad: null args");
  const aread_grouped_wgrad_args& a = *args;
  AREAD_REQUIRE(a.m >= 0 && a.n > 0 && a.k > 0, "grouped_wgrad: bad shape m=%lld n=%d k=%d", (long long)a.m, a.n, a.k);
  AREAD_REQUIRE(a.groups > 0 && a.groups <= kMaxGroups, "grouped_wgrad: groups %d not in [1, %d]", a.groups,
                kMaxGroups);
  AREAD_REQUIRE(a.dz && a.a && a.dw && a.workspace, "grouped_wgrad: null pointer");
  AREAD_REQUIRE(a.ldz % 8 == 0 && a.lda % 8 == 0, "grouped_wgrad: ldz/lda must be multiples of 8 bf16 (16 bytes)");
  AREAD_REQUIRE(reinterpret_cast<uintptr_t>(a.dz) % 16 == 0 && reinterpret_cast<uintptr_t>(a.a) % 16 == 0,
                "grouped_wgrad: operands must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  WgradParams p{};
  int bn;
  wgrad_plan(a, &p, &bn);
  if (p.n_active == 0) return AREAD_OK;
  const int64_t group_elems = static_cast<int64_t>(a.n) * a.k;
  if (a.m == 0) {  // empty batch: the gradient of the active groups is zero
    for (int gi = 0; gi < p.n_active; ++gi)
      AREAD_CUDA(cudaMemsetAsync(a.dw + p.group_ids[gi] * group_elems, 0, group_elems * 4, stream));
    return AREAD_OK;
  }
  const size_t need = static_cast<size_t>(p.n_split) * p.n_active * group_elems * 4;
  if (need > a.workspace_bytes)
    return fail(AREAD_ERR_WORKSPACE, "grouped_wgrad: workspace %zu < %zu bytes", a.workspace_bytes, need);
  p.partial = static_cast<float*>(a.workspace);

  AREAD_REQUIRE((a.dz_lo != nullptr) == (a.a_lo != nullptr), "grouped_wgrad: dz_lo and a_lo go together");
  p.n_seg = a.dz_lo != nullptr ? 3 : 1;
  CUtensorMap mdz, ma, mdz_lo, ma_lo;
  const int64_t a_cols = a.a_group_cols == 0 ? a.k : static_cast<int64_t>(a.a_group_cols) * (a.groups - 1) + a.k;
  if (int rc = make_map(&mdz, a.dz, a.m, static_cast<int64_t>(a.groups) * a.n, a.ldz, BK, 64)) return rc;
  if (int rc = make_map(&ma, a.a, a.m, a_cols, a.lda, BK, 64)) return rc;
  mdz_lo = mdz;
  ma_lo = ma;
  if (p.n_seg == 3) {
    AREAD_REQUIRE(reinterpret_cast<uintptr_t>(a.dz_lo) % 16 == 0 && reinterpret_cast<uintptr_t>(a.a_lo) % 16 == 0,
                  "grouped_wgrad: lo operands must be 16-byte aligned");
    if (int rc = make_map(&mdz_lo, a.dz_lo, a.m, static_cast<int64_t>(a.groups) * a.n, a.ldz, BK, 64)) return rc;
    if (int rc = make_map(&ma_lo, a.a_lo, a.m, a_cols, a.lda, BK, 64)) return rc;
  }
  if (int rc = (bn == 128 ? launch_wgrad<128>(mdz, ma, mdz_lo, ma_lo, p, stream)
                          : launch_wgrad<64>(mdz, ma, mdz_lo, ma_lo, p, stream)))
    return rc;
  const int64_t total = static_cast<int64_t>(p.n_active) * group_elems;
  int64_t grid = (total + 255) / 256;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  AREAD_LAUNCH(wgrad_reduce_kernel, static_cast<unsigned>(grid), 256, 0, stream, p.partial, a.dw, p.n_split,
               p.n_active, group_elems, p);
  return AREAD_OK;
}
