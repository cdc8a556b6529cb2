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
  unsigned char group_ids[kMaxGroups];
};

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreOffset = kStages * kStageBytes;           // epilogue staging: 4 warps x 2 x [32][32] fp32
  static constexpr int kStoreBytes = 4 * 2 * 32 * 128;
  static constexpr int kBarrierOffset = kStoreOffset + kStoreBytes;
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

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
grouped_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_lo,
                      const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
        const int b_row0 = c.g * p.n + c.n_t * BN;
        for (int seg = 0; seg < p.n_seg; ++seg) {
          const CUtensorMap* ma = seg < 2 ? &map_a : &map_a_lo;
          const CUtensorMap* mb = seg == 1 ? &map_b_lo : &map_b;
          for (int kb = 0; kb < p.n_k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            ptx::tma_load_2d(sa, ma, &full_bar[stage], a_col0 + kb * BK, c.m_t * BM);
            ptx::tma_load_2d(sb, mb, &full_bar[stage], kb * BK, b_row0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
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
            const uint64_t db = ptx::umma_desc_k_sw128(sb + kk * UMMA_K * 2, 8 * 128);
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
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(p, tile);
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int64_t row = static_cast<int64_t>(c.m_t) * BM + quad * 32 + lane;
      const int n0 = c.n_t * BN;  // column inside the group
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
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<unsigned*>(&h0);
                pk.y = *reinterpret_cast<unsigned*>(&h1);
                pk.z = *reinterpret_cast<unsigned*>(&h2);
                pk.w = *reinterpret_cast<unsigned*>(&h3);
                *reinterpret_cast<uint4*>(dst + j) = pk;
              }
            } else {
              for (int j = 0; j < 32 && col_in_group + j < p.n; ++j) dst[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && lane == 0) ptx::tma_store_wait_read<0>();   // shared memory must outlive the last stores
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
             int box_cols = BK) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return AREAD_OK;
}

// fp32 row-major [rows, cols] output, box = [32 rows, 32 cols] (one 128-byte swizzle row per matrix row)
int make_store_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AREAD_ERR_CUDA, "cuTensorMapEncodeTiled (store) failed with CUresult %d", (int)r);
  return AREAD_OK;
}

template <int BN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma_lo, const CUtensorMap& mb_lo,
                const CUtensorMap& mc, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    AREAD_CUDA(cudaFuncSetAttribute(grouped_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kTotal));
    configured = true;
  }
  const unsigned grid = static_cast<unsigned>(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
  AREAD_LAUNCH((grouped_linear_kernel<BN>), grid, kGemmThreads, L::kTotal, stream, ma, mb, ma_lo, mb_lo, mc, p);
  return AREAD_OK;
}


template <int BN>
int launch_wgrad(const CUtensorMap& mdz, const CUtensorMap& ma, const CUtensorMap& mdz_lo, const CUtensorMap& ma_lo,
                 const WgradParams& p, cudaStream_t stream) {
  constexpr int kSmem = kStages * (2 + BN / 64) * (BK * 64 * 2) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    AREAD_CUDA(cudaFuncSetAttribute(grouped_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  const unsigned grid = static_cast<unsigned>(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs);
  AREAD_LAUNCH((grouped_wgrad_kernel<BN>), grid, kGemmThreads, kSmem, stream, mdz, ma, mdz_lo, ma_lo, p);
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
  p.n_seg = a.a_lo != nullptr ? 3 : 1;
  CUtensorMap ma, mb, ma_lo, mb_lo;
  const int64_t a_cols = a.a_group_cols == 0 ? a.k : static_cast<int64_t>(a.a_group_cols) * (a.groups - 1) + a.k;
  if (int rc = make_map(&ma, a.a, a.m, a_cols, a.lda, BM)) return rc;
  if (int rc = make_map(&mb, a.b, static_cast<int64_t>(a.groups) * a.n, a.k, a.ldb, bn)) return rc;
  ma_lo = ma;
  mb_lo = mb;
  if (p.n_seg == 3) {
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
  return bn == 128 ? launch_gemm<128>(ma, mb, ma_lo, mb_lo, mc, p, stream)
                   : launch_gemm<64>(ma, mb, ma_lo, mb_lo, mc, p, stream);
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
