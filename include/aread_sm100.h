/*
 * aread_sm100.h -- C ABI of libaread_sm100.so: the AREAD hot path as hand-written sm_100a CUDA.
 *
 * The reference (Chrissie-Law/AREAD-Multi-Domain-Recommendation) is pure PyTorch and has no FFI of
 * its own; each entry point below names the reference function (file:line under the reference
 * tree) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless a
 *     field says "host".  The library never allocates device memory and keeps no global state
 *     besides a thread-local error string.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     and is safe to capture in a CUDA graph unless stated otherwise.
 *   - return value: AREAD_OK or a negative aread_status; aread_last_error() describes the failure.
 *   - workspaces: `aread_<op>_workspace_bytes()` gives the scratch size; pass a buffer at least
 *     that large, 256-byte aligned.
 *   - fp32 tensors are row-major and contiguous; "rows" are samples, never padded.
 */
#ifndef AREAD_SM100_H
#define AREAD_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AREAD_API __attribute__((visibility("default")))
#else
#define AREAD_API
#endif

typedef void* aread_stream_t; /* cudaStream_t */

typedef enum aread_status {
  AREAD_OK = 0,
  AREAD_ERR_INVALID = -1, /* bad argument (null pointer, unsupported size, misalignment)            */
  AREAD_ERR_INDEX = -2,   /* an id + field offset fell outside [0, n_rows): torch raises IndexError */
  AREAD_ERR_CUDA = -3,    /* a CUDA runtime call failed                                            */
  AREAD_ERR_WORKSPACE = -4 /* workspace too small                                                   */
} aread_status;

/* Human-readable description of the last failure on the calling thread ("" if none). */
AREAD_API const char* aread_last_error(void);
/* ABI version of the loaded library (bumped whenever a struct below changes). */
AREAD_API int aread_abi_version(void);
/* Number of kernel launches issued through this library by the calling process so far. */
AREAD_API uint64_t aread_launch_count(void);
/* A host that replays library launches recorded in a CUDA graph reports them here (the counter itself only
 * sees launches made through the entry points). */
AREAD_API void aread_launch_count_add(uint64_t n);

/* ------------------------------------------------------------------------------------------------
 * Multi-field embedding lookup.
 * Replaces FeaturesEmbedding.forward (model/layer.py:160-183): idx = x + offsets (:165), table
 * gather (:166), mean/sum pooling of multi-hot columns over seq_maxlen positions, padding positions
 * included (:168-176), concat [one-hot fields | pooled fields] (:178).
 *
 * The layout of the lookup is described once by a "plan" living in device memory:
 *   col_offset[c]            row offset added to column c of x              (layer.py:151-157)
 *   field_src[f * max_src+k] k-th input column feeding output field f (k < field_nsrc[f])
 *   field_nsrc[f]            1 for a one-hot field, seq_maxlen for a pooled field
 *   field_div[f]             divisor applied after the in-order sum (1 or seq_maxlen for 'mean')
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_embed_plan {
  int32_t n_cols;            /* columns of x                                   */
  int32_t n_fields;          /* output fields (output_dim0, layer.py:141-143)  */
  int32_t max_src;           /* stride of field_src                            */
  int32_t embed_dim;         /* D; must be a multiple of 4                     */
  int64_t n_rows;            /* rows of the table (sum of the one-hot dims)    */
  const int32_t* col_offset; /* [n_cols]                                       */
  const int32_t* field_src;  /* [n_fields * max_src]                           */
  const int32_t* field_nsrc; /* [n_fields]                                     */
  const float* field_div;    /* [n_fields]                                     */
} aread_embed_plan;

typedef struct aread_gather_args {
  aread_embed_plan plan;
  int64_t batch;       /* B                                                                       */
  const int32_t* x;    /* [B, n_cols] ids (run.py:274 stores them as int32)                        */
  const float* table;  /* [n_rows, D] fp32                                                         */
  float* out;          /* [B, n_fields, D] fp32, bit-exact copy / in-order pooled sum              */
  uint16_t* out_bf16;  /* optional [B, n_fields * D] bf16 (round-to-nearest-even) copy, or NULL    */
  uint16_t* out_bf16_lo; /* optional [B, n_fields * D]: bf16(out - bf16(out)), the split residual  */
  int32_t* status;     /* [2] device ints: status[0] != 0 after an out-of-range id, status[1] = the
                          offending row index.  Zero it before the first call.  Rows that are out
                          of range produce zeros.                                                 */
  /* Row-sharded table over 2^shard_shift GPUs of one NVSwitch box (0 = not sharded, `table` is used).
     Global row r lives on GPU r mod 2^shard_shift at local row r >> shard_shift; shards[g] is the
     device pointer of GPU g's shard [shard_rows, D], peer-mapped into this process, so the gather
     reads remote rows straight over NVLink: the lookup IS the all-to-all of rows.                 */
  int32_t shard_shift;
  const float* const* shards; /* device array [2^shard_shift] of device pointers                  */
} aread_gather_args;

AREAD_API int aread_gather_fwd(const aread_gather_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient of the lookup: deterministic sort-by-row segmented scatter-add.
 * Replaces the autograd of model/layer.py:166 (aten::embedding_dense_backward, dense because
 * sparse=False at layer.py:150) including the 1/seq_maxlen factor of the mean pooling.
 *
 * Order of summation (documented so that it can be reproduced bit for bit, oracle/embedding_np.py
 * scatter_bwd_tiled): lookups are stably sorted by table row; gradient rows of mean-pooled columns
 * are scaled by fl32(1 / seq_maxlen); the sorted list is covered by an aligned 32-ary tree whose
 * level-l blocks hold AREAD_SCATTER_TILE^l consecutive entries.  A row's entries inside one
 * level-1 block (tile) are summed left to right; its sum inside a level-l block is the left to
 * right sum of its sums inside that block's children.  A row whose entries all fall in one tile
 * therefore reproduces the reference (sequential) order exactly.
 * ---------------------------------------------------------------------------------------------- */
#define AREAD_SCATTER_TILE 32

typedef struct aread_scatter_args {
  aread_embed_plan plan;
  int64_t batch;
  const int32_t* x;     /* [B, n_cols]                                                          */
  const float* d_out;   /* [B, n_fields, D] gradient w.r.t. the gather output                   */
  float* d_table;       /* [n_rows, D]; rows that were looked up are OVERWRITTEN with their sum,
                           other rows are left untouched (zero-fill it first for the dense
                           gradient of the reference, see zero_fill)                            */
  int32_t zero_fill;    /* != 0: clear d_table on `stream` before scattering                    */
  void* workspace;      /* aread_scatter_workspace_bytes(B * n_cols, D) bytes                   */
  size_t workspace_bytes;
  int32_t* sorted_rows; /* optional out [B * n_cols]: table rows in sorted order, or NULL       */
  int32_t* sorted_pos;  /* optional out [B * n_cols]: flattened (b, c) position of each, or NULL */
  /* Owner-major output for a row-sharded table: with shard_shift > 0, d_table is
     [2^shard_shift, shard_rows, D] and global row r is written at (r mod 2^shard_shift, r >> shard_shift),
     ready for a reduce-scatter that hands every GPU the summed gradient of its own shard.       */
  int32_t shard_shift;
  int64_t shard_rows;
  /* Sparse exchange of the gradient of a row-sharded table: with peer_grads != NULL (device array [2^shard_shift] of
     device pointers), the reduced gradient of table row r is stored at peer_grads[r mod N] + (r >> shard_shift) * D --
     i.e. straight into the per-sender receive buffer of the rank that owns the row, over NVLink (P2P store).  Only
     rows this batch touched travel; d_table is not written (and zero_fill is ignored).  The owner sums its N receive
     buffers with aread_shard_grad_sum.                                                                           */
  float* const* peer_grads;
} aread_scatter_args;

AREAD_API size_t aread_scatter_workspace_bytes(int64_t n_lookups, int32_t embed_dim);
AREAD_API int aread_scatter_bwd(const aread_scatter_args* args, aread_stream_t stream);
/* Owner side of the sparse exchange: out[i] = scale * sum_s recv[s * n + i] (senders added in rank order:
 * deterministic); every element of recv that was non-zero is reset to zero, so the buffers are ready for the next
 * step without a dense memset.  recv: [n_senders, n] fp32, n % 4 == 0, 16-byte aligned.                            */
AREAD_API int aread_shard_grad_sum(float* recv, int32_t n_senders, int64_t n, float scale, float* out,
                                   aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Grouped Linear on the tensor cores (tcgen05 / TMEM, operands staged by TMA).
 * Replaces the nn.Linear calls inside MultiLayerPerceptron.forward (model/layer.py:210, 221-229)
 * for the four MMoE experts (model/aread.py:93-95, 150) -- group = expert -- and their gradients.
 *
 *   C[:, g*n : (g+1)*n] = A[:, g*a_group_cols : g*a_group_cols + k] . B[g*n : (g+1)*n, :]^T (+ bias)
 *
 * for every group g whose bit is set in group_mask; groups that are masked out are skipped entirely
 * (no tiles are scheduled for them and their output columns are left untouched).  bf16 operands,
 * fp32 accumulation; the result is stored as fp32 or as bf16 (round to nearest even).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_grouped_linear_args {
  int64_t m;            /* rows (samples)                                                          */
  int32_t n;            /* output width per group                                                  */
  int32_t k;            /* reduction length per group                                              */
  int32_t groups;       /* number of groups, <= 64                                                 */
  int32_t a_group_cols; /* group g reads A columns starting at g * a_group_cols; 0 = all groups
                           share the same A columns [0, k)                                         */
  uint64_t group_mask;  /* bit g set = compute group g                                             */
  const uint16_t* a;    /* bf16 [m, lda], 16-byte aligned, lda % 8 == 0                            */
  int64_t lda;
  const uint16_t* b;    /* bf16 [groups * n, ldb] (nn.Linear weight layout, groups stacked by rows) */
  int64_t ldb;
  const float* bias;    /* optional fp32 [groups * n]                                              */
  float* c_f32;         /* exactly one of c_f32 / c_bf16: [m, ldc]                                 */
  uint16_t* c_bf16;
  int64_t ldc;
  const uint16_t* a_lo; /* optional split-precision residuals (same shapes / strides as a and b):  */
  const uint16_t* b_lo; /* x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  When given, the product
                           is a.b + a.b_lo + a_lo.b (three tensor-core passes into one accumulator):
                           fp32-grade results (~2^-16 relative) instead of bf16 operand rounding   */
  int32_t lo_lo;        /* 1: also add a_lo.b_lo (a fourth pass): the full product of the split operands  */
} aread_grouped_linear_args;

AREAD_API int aread_grouped_linear_bf16(const aread_grouped_linear_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Expert Linear layers with the BatchNorm bookkeeping fused into the GEMM epilogue (bf16 experts).
 * Same tcgen05 / TMEM / TMA pipeline as aread_grouped_linear_bf16; the accumulator leaves TMEM once and the
 * pre-activation is stored ONCE, as bf16 -- the fp32 `z` round trips of Linear -> BatchNorm1d -> ReLU -> Dropout
 * (model/layer.py:209-215, 221-229) and of their backward are gone.
 *
 *   acc[:, g*n : (g+1)*n] = A[:, g*a_group_cols : +k] . B_g        (fp32 accumulation in TMEM)
 *
 * B_g: b_is_k_by_n == 0 -> rows g*n.. of b [groups*n, ldb], k contiguous (the nn.Linear weight, forward);
 *      b_is_k_by_n == 1 -> rows g*k.. of b [groups*k, ldb], n contiguous (the SAME weight array read as an
 *                          MN-major operand: data gradient dA = dZ . W without a transposed copy).
 * Epilogues:
 *   AREAD_EPI_PLAIN   c = acc (+ bias), fp32 or bf16
 *   AREAD_EPI_STATS   c_bf16 = bf16(acc)  [bias-free: BatchNorm removes it]; partial[t][0][col] = sum_rows acc,
 *                     partial[t][1][col] = sum_rows acc^2 over the rows of 128-row tile t (fixed order)
 *   AREAD_EPI_ACT     c_bf16 = bf16(max(acc * scale[col] + shift[col], 0))   (eval mode: BatchNorm folded)
 *   AREAD_EPI_BF16    c_bf16 = bf16(acc) through the same 64-column TMA-store path (n % 64 == 0), nothing else
 *   AREAD_EPI_BN_BWD  acc = gradient w.r.t. the ACTIVATED output of the layer below;  with y = z*scale+shift,
 *                     dy = (y > 0 && keep) ? acc / (1 - p) : 0;  c_bf16 = bf16(dy);
 *                     partial[t][0][col] = sum dy, partial[t][1][col] = sum dy * (z - mean) * rstd
 * partial: fp32 [aread_expert_gemm_partials(m)][2][groups * n].
 * ---------------------------------------------------------------------------------------------- */
enum { AREAD_EPI_PLAIN = 0, AREAD_EPI_STATS = 1, AREAD_EPI_ACT = 2, AREAD_EPI_BN_BWD = 3, AREAD_EPI_BF16 = 4 };

typedef struct aread_expert_gemm_args {
  int64_t m;
  int32_t n, k, groups, a_group_cols;
  const uint16_t* a;      /* bf16 [m, lda] */
  int64_t lda;
  const uint16_t* b;      /* bf16 weights, see b_is_k_by_n */
  int64_t ldb;
  int32_t b_is_k_by_n;
  int32_t epilogue;       /* AREAD_EPI_*                                              */
  const float* bias;      /* PLAIN only, optional fp32 [groups * n]                   */
  float* c_f32;           /* PLAIN only (or c_bf16)                                   */
  uint16_t* c_bf16;       /* [m, ldc] bf16                                            */
  int64_t ldc;
  float* partial;         /* STATS / BN_BWD                                           */
  const float* scale;     /* ACT / BN_BWD: fp32 [groups * n]                          */
  const float* shift;
  const float* mean;      /* BN_BWD                                                   */
  const float* rstd;
  const uint16_t* z;      /* BN_BWD: bf16 [m, ldz] pre-activation of the layer below  */
  int64_t ldz;
  float dropout_p;        /* BN_BWD: dropout of the layer below (0: none)             */
  uint32_t salt;
  uint64_t seed;
  const uint64_t* seed_ptr;
} aread_expert_gemm_args;

AREAD_API int32_t aread_expert_gemm_partials(int64_t m);
AREAD_API int aread_expert_gemm(const aread_expert_gemm_args* args, aread_stream_t stream);

/* Column statistics of a STATS launch -> what BatchNorm1d needs (model/layer.py:212, torch semantics: batch mean /
 * biased variance in training, running statistics with momentum and the unbiased variance).  `mean` is relative to
 * the bias-free accumulator that was stored; the running mean gets mean + bias.  training == 0: running statistics
 * (shift absorbs the bias); bn_skip (batch of one, layer.py:226): scale = 1, shift = bias.                        */
typedef struct aread_expert_bn_finalize_args {
  int64_t m;
  int32_t width, n_partial, training, bn_skip;
  float momentum, eps;
  const float* partial;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean;            /* out, fp32 [width] each */
  float* rstd;
  float* scale;
  float* shift;
} aread_expert_bn_finalize_args;

AREAD_API int aread_expert_bn_finalize(const aread_expert_bn_finalize_args* args, aread_stream_t stream);

/* Column sums of a BN_BWD launch -> d_gamma = sum dy*xhat, d_beta = sum dy, d_bias = 0 (BatchNorm removes the column
 * mean; with bn_skip: d_bias = sum dy, the others 0) and coef = [sum dy / m | sum dy*xhat / m].                    */
typedef struct aread_expert_bn_bwd_finalize_args {
  int64_t m;
  int32_t width, n_partial, bn_skip;
  const float* partial;
  float* d_gamma;
  float* d_beta;
  float* d_bias;
  float* coef;            /* out, fp32 [2, width] */
} aread_expert_bn_bwd_finalize_args;

AREAD_API int aread_expert_bn_bwd_finalize(const aread_expert_bn_bwd_finalize_args* args, aread_stream_t stream);

/* out = bf16(dropout(relu(z * scale + shift))) from the bf16 pre-activation (forward), or, with dy != NULL,
 * dz = bf16(scale * (dy - coef[0] - xhat * coef[1])) (bn_skip: dz = dy), xhat = (z - mean) * rstd (backward).
 * HBM-bound passes: a thread owns 8 columns (per-column constants in registers) and walks rows.                   */
typedef struct aread_bn16_args {
  int64_t m;
  int32_t width, bn_skip;
  const uint16_t* z;      /* bf16 [m, ldz]                      */
  int64_t ldz;
  const float* scale;
  const float* shift;
  float dropout_p;        /* forward: 0 in eval mode            */
  uint32_t salt;
  uint64_t seed;
  const uint64_t* seed_ptr;
  uint16_t* out;          /* bf16 [m, ldo]: h (forward) / dz (backward) */
  int64_t ldo;
  const uint16_t* dy;     /* backward: bf16 [m, ldd], masks already applied; NULL: forward */
  int64_t ldd;
  const float* mean;      /* backward */
  const float* rstd;
  const float* coef;      /* backward: fp32 [2, width] */
  int32_t dy_is_raw;      /* backward: 1 = `dy` is the gradient w.r.t. the ACTIVATED output (a plain data-gradient GEMM
                             wrote it); the ReLU / dropout mask is applied here: read from pass_bits when given, else
                             rebuilt from z, scale, shift, dropout_p, seed                                          */
  uint8_t* pass_bits;     /* optional [m, width / 8]: bit j of byte (row, g) = element (row, 8 g + j) passed ReLU and
                             dropout.  Written by the forward, read by the backward passes                          */
  float keep_scale_bwd;   /* set by the library                                                                     */
} aread_bn16_args;

AREAD_API int aread_bn16(const aread_bn16_args* args, aread_stream_t stream);
/* Reduction half of the backward on the bf16 tensors: partial[c][0][col] = sum over CTA c's rows of dy,
 * partial[c][1][col] = sum of dy * xhat (rows in order), c < aread_bn16_partials(m, width); feed them to
 * aread_expert_bn_bwd_finalize.                                                                        */
AREAD_API int32_t aread_bn16_partials(int64_t m, int32_t width);
AREAD_API int aread_bn16_bwd_stats(const aread_bn16_args* args, float* partial, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weight gradient of the grouped Linear: dW_g[j, i] = sum_b dZ[b, g*n + j] * A[b, g*a_group_cols + i].
 * Replaces the autograd of the nn.Linear weights (model/layer.py:210) of the experts.  Both operands
 * are read in their natural [samples, features] layout (MN-major tcgen05 operands); the sample
 * range is split over CTAs and the fp32 partials are summed in a fixed order (deterministic).
 * Rows of dw that belong to masked-out groups are left untouched.
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_grouped_wgrad_args {
  int64_t m;            /* samples                                                  */
  int32_t n;            /* output features per group                                */
  int32_t k;            /* input features per group                                 */
  int32_t groups;
  int32_t a_group_cols; /* as in aread_grouped_linear_args                          */
  uint64_t group_mask;
  const uint16_t* dz;   /* bf16 [m, ldz]: gradient w.r.t. the Linear output         */
  int64_t ldz;
  const uint16_t* a;    /* bf16 [m, lda]: the Linear input                          */
  int64_t lda;
  float* dw;            /* fp32 [groups * n, k] contiguous                          */
  void* workspace;      /* aread_grouped_wgrad_workspace_bytes(args) bytes          */
  size_t workspace_bytes;
  const uint16_t* dz_lo; /* optional split-precision residuals, see aread_grouped_linear_args */
  const uint16_t* a_lo;
} aread_grouped_wgrad_args;

AREAD_API size_t aread_grouped_wgrad_workspace_bytes(const aread_grouped_wgrad_args* args);
AREAD_API int aread_grouped_wgrad_bf16(const aread_grouped_wgrad_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm1d + ReLU + Dropout after a Linear, forward and backward.
 * Replaces the BatchNorm1d / ReLU / Dropout slots of MultiLayerPerceptron.forward
 * (model/layer.py:211-215, 225-228): training uses batch mean / biased variance (eps 1e-5) and
 * updates running_mean / running_var with momentum 0.1 and the unbiased variance; eval uses the
 * running statistics; bn_skip = 1 is the batch-of-one rule (layer.py:226, BatchNorm not applied).
 * Dropout keeps element (row, col) iff hash(seed, salt, row * width + col) >= p * 2^32 and scales by
 * 1 / (1 - p); aread_dropout_mask exposes the same stream so tests can feed it to the oracle.
 * Column sums are reduced per CTA over fixed row ranges and added in CTA order (deterministic).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_bn_act_args {
  int64_t m;              /* rows                                                                  */
  int32_t width;          /* columns (all groups of the layer side by side), <= 2048               */
  int32_t training;       /* 1: batch statistics + running update + dropout, 0: running statistics */
  int32_t bn_skip;        /* 1: BatchNorm is the identity (batch of one)                           */
  float momentum, eps, dropout_p;
  uint64_t seed;          /* dropout stream                                                        */
  uint32_t salt;          /* per-layer constant mixed into the stream                              */
  const float* z;         /* fp32 [m, ldz]: Linear output                                          */
  int64_t ldz;
  const float* gamma;     /* [width] BatchNorm weight                                              */
  const float* beta;      /* [width] BatchNorm bias                                                */
  float* running_mean;    /* [width], updated in place when training                               */
  float* running_var;
  float* mean;            /* out [width]: statistics used (saved for the backward)                 */
  float* rstd;
  float* scale;           /* out [width]: gamma * rstd                                             */
  float* shift;           /* out [width]: beta - mean * scale                                      */
  float* out_f32;         /* optional out [m, ldo]: dropout(relu(bn(z)))                           */
  uint16_t* out_bf16;     /* optional out [m, ldo] as bf16                                         */
  int64_t ldo;
  void* workspace;        /* aread_bn_workspace_bytes(width)                                       */
  size_t workspace_bytes;
  uint16_t* out_bf16_lo;  /* optional out [m, ldo]: bf16(value - bf16(value)), the split residual  */
  const uint64_t* seed_ptr; /* optional device scalar: when set, the dropout seed is read from it at run time   */
                            /* (CUDA-graph replay) and `seed` is ignored                                        */
} aread_bn_act_args;

typedef struct aread_bn_act_bwd_args {
  int64_t m;
  int32_t width;
  int32_t bn_skip;
  float dropout_p;        /* 0 when the forward ran without dropout                                */
  uint32_t salt;
  uint64_t seed;
  const float* z;         /* fp32 [m, ldz]: the forward's Linear output                            */
  int64_t ldz;
  const float* d_out;     /* fp32 [m, ldd]: gradient w.r.t. dropout(relu(bn(z)))                   */
  int64_t ldd;
  const float* mean;      /* saved by the forward                                                  */
  const float* rstd;
  const float* scale;
  const float* shift;
  float* d_gamma;         /* optional out [width]                                                  */
  float* d_beta;          /* optional out [width]                                                  */
  float* d_bias;          /* optional out [width]: gradient of the Linear bias (exactly 0 under BatchNorm) */
  float* dz_f32;          /* optional out [m, ldo]: gradient w.r.t. z                              */
  uint16_t* dz_bf16;      /* optional out [m, ldo] as bf16                                         */
  int64_t ldo;
  void* workspace;
  size_t workspace_bytes;
  uint16_t* dz_bf16_lo;   /* optional out [m, ldo]: the split residual of dz                       */
  const uint64_t* seed_ptr; /* optional device scalar: when set, the dropout seed is read from it at run time   */
                            /* (CUDA-graph replay) and `seed` is ignored                                        */
  const uint16_t* z_bf16; /* optional: the pre-activation as bf16 [m, ldz] (aread_expert_gemm STATS); `z` is  */
                          /* then ignored                                                                     */
} aread_bn_act_bwd_args;

AREAD_API size_t aread_bn_workspace_bytes(int32_t width);
AREAD_API int aread_bn_act_fwd(const aread_bn_act_args* args, aread_stream_t stream);
AREAD_API int aread_bn_act_bwd(const aread_bn_act_bwd_args* args, aread_stream_t stream);
AREAD_API int aread_dropout_mask(uint64_t seed, uint32_t salt, int64_t n, float p, uint8_t* out, aread_stream_t stream);

/* Activation pass of aread_bn_act_fwd alone: out = dropout(relu(z * scale + shift)) with scale / shift given. */
AREAD_API int aread_bn_act_apply(const aread_bn_act_args* args, aread_stream_t stream);
/* Reduction half of aread_bn_act_bwd alone: d_gamma / d_beta / d_bias and
 * coef[0:width] = sum(dy) / m, coef[width:2*width] = sum(dy * xhat) / m. */
AREAD_API int aread_bn_bwd_coef(const aread_bn_act_bwd_args* args, float* coef, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * One HEI tower layer for all towers that run under the mask, fused with its surroundings
 * (model/layer.py:221-229 inside model/aread.py:297-321).  `src` is [m, groups * k]: either a plain
 * activation (src_scale == NULL) or the PREVIOUS layer's pre-activation z, in which case
 * dropout(relu(src * src_scale + src_shift)) is formed on the fly with the dropout stream
 * (seed, src_salt) -- the activation between two layers never exists in memory.
 *
 * forward:  z = act(src) W^T + bias, plus this layer's BatchNorm statistics (mean, rstd, scale, shift,
 *           running update) in the same pass over z.
 * backward: given d_out (gradient w.r.t. this layer's activation) and coef (aread_bn_bwd_coef of
 *           (z, d_out), or the src_coef a previous call produced): d_w = dz^T act(src),
 *           d_in = dz W (gradient w.r.t. act(src)), and when src is a pre-activation the coefficients
 *           and parameter gradients of ITS BatchNorm backward (src_coef, src_d_gamma / _beta / _bias).
 * k, n <= 64 (aread_hei_layer_supported); wider layers go through aread_tower_linear & co.
 *
 * Two implementations sit behind these entry points: the tcgen05 kernels (csrc/hei_tc.cu; k in {16, 32, 64},
 * 64 n / k in {32, 64}, m >= 512) and the CUDA-core kernels (csrc/hei.cu; everything else).  aread_hei_set_path
 * overrides the choice for the process: 1 = tensor cores where the shape allows, 0 = CUDA cores, -1 = the
 * environment's default (AREAD_HEI_TC, AREAD_HEI_TC_BWD; both on).  aread_hei_layer_path reports what a call of
 * that shape would run: bit 0 = forward on tensor cores, bit 1 = backward on tensor cores.
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_hei_layer_fwd_args {
  int64_t m;
  int32_t groups, k, n;
  int32_t training, bn_skip;
  float momentum, eps;
  const float* src;          /* [m, ld_src], group g at columns g*k .. g*k + k                */
  int64_t ld_src;
  const float* src_scale;    /* [groups * k] or NULL                                          */
  const float* src_shift;
  float src_p;               /* dropout of the src activation (applied when training)         */
  uint32_t src_salt;
  uint64_t seed;
  const float* weight;       /* [groups, n, k]                                                */
  const float* bias;         /* [groups, n] or NULL                                           */
  const float* gamma;        /* [groups * n]                                                  */
  const float* beta;
  float* running_mean;       /* [groups * n], updated in place when training                  */
  float* running_var;
  float* z;                  /* out [m, groups * n]                                           */
  float* mean;               /* out [groups * n] each                                         */
  float* rstd;
  float* scale;
  float* shift;
  void* workspace;           /* aread_hei_layer_workspace_bytes(m, groups, k, n)              */
  size_t workspace_bytes;
  const uint64_t* seed_ptr; /* optional device scalar: when set, the dropout seed is read from it at run time   */
                            /* (CUDA-graph replay) and `seed` is ignored                                        */
} aread_hei_layer_fwd_args;

typedef struct aread_hei_layer_bwd_args {
  int64_t m;
  int32_t groups, k, n;
  int32_t bn_skip;
  float p;                   /* dropout of THIS layer's activation in the forward (0 in eval)  */
  uint32_t salt;
  uint64_t seed;
  const float* z;            /* [m, groups * n]                                               */
  const float* d_out;        /* [m, groups * n]                                               */
  const float* mean;         /* this layer's saved statistics, [groups * n] each              */
  const float* rstd;
  const float* scale;
  const float* shift;
  const float* coef;         /* [2, groups * n]                                               */
  const float* src;          /* as in the forward                                             */
  int64_t ld_src;
  const float* src_scale;
  const float* src_shift;
  const float* src_mean;
  const float* src_rstd;
  float src_p;
  uint32_t src_salt;
  const float* weight;       /* [groups, n, k]                                                */
  float* d_in;               /* out [m, groups * k] or NULL                                   */
  float* d_w;                /* out [groups, n, k]                                            */
  float* src_coef;           /* out [2, groups * k]      (src_scale != NULL)                  */
  float* src_d_gamma;        /* out [groups * k] each, optional                               */
  float* src_d_beta;
  float* src_d_bias;
  void* workspace;
  size_t workspace_bytes;
  const uint64_t* seed_ptr; /* optional device scalar: when set, the dropout seed is read from it at run time   */
                            /* (CUDA-graph replay) and `seed` is ignored                                        */
} aread_hei_layer_bwd_args;

AREAD_API int aread_hei_layer_supported(int32_t groups, int32_t k, int32_t n);
AREAD_API void aread_hei_set_path(int32_t tensor_cores_fwd, int32_t tensor_cores_bwd);
AREAD_API int aread_hei_layer_path(int64_t m, int32_t groups, int32_t k, int32_t n);
AREAD_API size_t aread_hei_layer_workspace_bytes(int64_t m, int32_t groups, int32_t k, int32_t n);
AREAD_API int aread_hei_layer_fwd(const aread_hei_layer_fwd_args* args, aread_stream_t stream);
AREAD_API int aread_hei_layer_bwd(const aread_hei_layer_bwd_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * MMoE mixture fused with the last expert layer's BatchNorm/ReLU/Dropout.
 * Replaces model/aread.py:150-153: h_e = expert_e(x) (here: dropout(relu(z_e * scale + shift))),
 * out[b, g, :] = sum_e gate[b, g, e] * h[b, e, :].  Forward when d_out == NULL, else backward:
 * d_h[b, e, :] = sum_g gate[b, g, e] * d_out[b, g, :], d_gate[b, g, e] = <d_out[b, g, :], h[b, e, :]>.
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_mmoe_mix_args {
  int64_t m;
  int32_t width;          /* expert output width (expert_dims[-1])                                 */
  int32_t n_expert;       /* <= 16                                                                 */
  int32_t n_gate;         /* gates = level-0 towers, <= 8                                          */
  float dropout_p;
  uint64_t seed;
  uint32_t salt;
  const float* z;         /* fp32 [m, ldz]: last expert Linear output, experts side by side        */
  int64_t ldz;
  const float* scale;     /* [n_expert * width] folded BatchNorm                                   */
  const float* shift;
  const float* gate;      /* fp32 [m, n_gate, n_expert] softmax gates                              */
  float* out;             /* forward out [m, n_gate, width]                                        */
  const float* d_out;     /* backward in [m, n_gate, width]                                        */
  float* d_h;             /* backward out [m, n_expert * width]                                    */
  float* d_gate;          /* backward out [m, n_gate, n_expert]                                    */
  const uint64_t* seed_ptr; /* optional device scalar replacing `seed` at run time (CUDA-graph replay)   */
  const uint16_t* z_bf16; /* optional: z as bf16 [m, ldz]; `z` is then ignored                             */
} aread_mmoe_mix_args;

AREAD_API int aread_mmoe_mix(const aread_mmoe_mix_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row pass: every dot product of the flattened embedding row with a parameter vector, in fp32.
 * Replaces FeaturesLinear.forward (model/layer.py:122-126), CrossNetwork.forward (layer.py:529-537),
 * the MMoE gates Linear + Softmax (model/aread.py:96-99, 152) and the cross-network part of the
 * towers_linear heads (aread.py:119-120, 307), forward and backward.
 *
 * P = X . W^T with the rows of W ordered [linear | gate g, expert e | cross layer k | head t]; then
 *   lin[b]        = P[b, 0] + offset[0]
 *   gate[b, g, :] = softmax_e(P[b, 1 + g*n_expert + e] + offset[..])
 *   alpha_0 = 1, s_k = alpha_k * P[b, cross k] + offset[cross k], alpha_{k+1} = alpha_k + s_k
 *   head[b, t]    = alpha_n * P[b, head t] + offset[head t]
 * where offset holds the row-independent constants (biases, w_k . beta_k, w_out_t[:E] . beta_n with
 * beta_k = b_0 + .. + b_{k-1}); the cross-network output itself is never materialised.
 * The backward returns d_x, d_w ([columns, e], reduced over rows in a fixed CTA order) and d_c
 * ([m, ldp]; its column sums are the gradients of `offset`).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_rowpass_args {
  int64_t m;
  int32_t e;             /* embed_output_dim                                                       */
  int32_t n_gate;        /* MMoE gates that are evaluated (active level-0 towers)                  */
  int32_t n_expert;
  int32_t n_cross;       /* cross layers                                                           */
  int32_t n_head;        /* output heads that are evaluated (active last-level towers)             */
  int32_t ldp;           /* row stride of p / d_p / d_c, >= 1 + n_gate*n_expert + n_cross + n_head  */
  const float* x;        /* fp32 [m, e]                                                            */
  const float* w;        /* fp32 [columns, e]                                                      */
  const float* offset;   /* fp32 [columns]                                                         */
  float* p;              /* [m, ldp]: forward out, backward in                                     */
  float* lin;            /* [m]                                                                    */
  float* gate;           /* [m, n_gate, n_expert]                                                  */
  float* alpha;          /* [m, n_cross + 1]                                                       */
  float* head;           /* [m, n_head]                                                            */
  const float* d_lin;    /* backward inputs (NULL = zero)                                          */
  const float* d_gate;
  const float* d_head;
  float* d_p;            /* backward scratch/out [m, ldp]                                          */
  float* d_c;            /* backward out [m, ldp]                                                  */
  float* d_x;            /* backward out [m, e] or NULL                                            */
  float* d_w;            /* backward out [columns, e]                                              */
  void* workspace;       /* aread_rowpass_workspace_bytes(m, e, columns)                           */
  size_t workspace_bytes;
  /* Tensor-core variant: with x == NULL the skinny products are NOT evaluated here.  Forward: `p` is an input
     (X . W^T from aread_grouped_linear_bf16 with split operands) and only the per-row epilogue runs.  Backward: only
     the per-row prologue runs (d_p, d_c) and, when dp16 is set, d_p is also written as split bf16 operands for the
     tensor-core products d_x = d_p . W and d_w = d_p^T . X:
       dp16[b, 0:W] = hi, dp16[b, W:2W] = hi, dp16[b, 2W:3W] = lo   (hi = bf16(d_p), lo = bf16(d_p - hi); W =
     dp16_width; columns beyond the last dot product are zero), row stride ld16 elements -- the layout that extends
     the expert layer-1 gradient [m, 4*256 | 3W] so that ONE data-gradient GEMM returns the sum of both paths.   */
  uint16_t* dp16;
  int64_t ld16;
  /* n_extra further dot products ride along behind the heads (the HEI gate logits, whose input is a slice of the
     embedding row): forward, p[:, cols .. cols + n_extra) is left to the caller; backward, d_p of those columns is
     an INPUT (written by aread_gate_mix) that is copied to d_c and into the split.  dp16_width = columns of each
     third of the split, a multiple of 32 >= cols + n_extra (0 = 32).                                           */
  int32_t n_extra;
  int32_t dp16_width;
  /* backward, optional: the column sums of d_c over the rows, [ldp] (what the caller needs of d_c: the gradients of the
     additive constants).  The prologue kernel adds its tiles up per CTA in row order into d_c_partial
     (aread_rowpass_prologue_ctas(m) * ldp floats) and a second launch adds the CTAs in order.  With d_c_sum set, d_c
     may be NULL and is then never written.                                                                         */
  float* d_c_sum;
  float* d_c_partial;
} aread_rowpass_args;

AREAD_API size_t aread_rowpass_workspace_bytes(int64_t m, int32_t e, int32_t n_cols);
AREAD_API int32_t aread_rowpass_prologue_ctas(int64_t m);
AREAD_API int aread_rowpass_fwd(const aread_rowpass_args* args, aread_stream_t stream);
AREAD_API int aread_rowpass_bwd(const aread_rowpass_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * L2 regulariser over a list of fp32 tensors in one pass: out[0] = sum_i l2[i] * sum(w_i^2), and its
 * gradient grads[i] = 2 * l2[i] * g * w_i.  Replaces BaseModel.get_regularization_loss
 * (model/layer.py:96-112).  The concatenation of the tensors is cut into chunks of
 * aread_l2_reg_chunk() elements (a chunk never straddles two tensors); chunk_start[i] is the index
 * of tensor i's first chunk.  Partials are added in a fixed order (bit-reproducible).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_l2_reg_args {
  int32_t n_tensors;
  int64_t n_chunks;              /* sum_i ceil(sizes[i] / chunk)                    */
  const float* const* tensors;   /* device array [n_tensors] of device pointers     */
  float* const* grads;           /* device array [n_tensors] (backward only)        */
  const int64_t* sizes;          /* device array [n_tensors]: elements per tensor   */
  const float* l2;               /* device array [n_tensors]                        */
  const int64_t* chunk_start;    /* device array [n_tensors]                        */
  float* out;                    /* device scalar (forward)                         */
  void* workspace;               /* n_chunks * 4 bytes (forward)                    */
  size_t workspace_bytes;
} aread_l2_reg_args;

AREAD_API int64_t aread_l2_reg_chunk(void);
AREAD_API int aread_l2_reg_fwd(const aread_l2_reg_args* args, aread_stream_t stream);
/* g_out: device scalar with the incoming gradient (NULL = 1) */
AREAD_API int aread_l2_reg_bwd(const aread_l2_reg_args* args, const float* g_out, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Grouped small Linear of the HEI towers, fp32 on the CUDA cores (16..64 wide layers).
 * Replaces the nn.Linear calls of the tower MLPs (model/layer.py:210 via model/aread.py:108-110,
 * 307, 319) and their gradients.  The caller passes only the towers that run under the HEMP mask
 * (compact group index g = 0 .. groups-1); pruned towers cost nothing.
 *
 *   out[b, g, j] = sum_i in[b, g, i] * M_g[i, j] (+ bias[g, j])
 * forward:       in = tower inputs, weight = nn.Linear weights [groups, out, in], weight_is_out_by_in = 1
 * data gradient: in = dz,           weight = the same tensor read as [groups, in = N, out = K],
 *                weight_is_out_by_in = 0 (so that out = dz . W)
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_tower_linear_args {
  int64_t m;
  int32_t groups;
  int32_t in_width;             /* <= 256                                                   */
  int32_t out_width;            /* <= 128                                                   */
  int32_t weight_is_out_by_in;  /* 1: weight[g] is [out_width, in_width]; 0: [in_width, out_width] */
  const float* in;              /* fp32 [m, ld_in], group g at columns g * in_group_stride  */
  int64_t ld_in;
  int64_t in_group_stride;      /* in_width for side-by-side groups, 0 = all groups read the same input */
  const float* weight;          /* fp32 [groups, ...] contiguous                            */
  const float* bias;            /* optional fp32 [groups * out_width]                       */
  float* out;                   /* fp32 [m, ld_out], group g at columns g * out_width       */
  int64_t ld_out;
} aread_tower_linear_args;

AREAD_API int aread_tower_linear(const aread_tower_linear_args* args, aread_stream_t stream);

/* d_w[g, n, k] = sum_b dz[b, g, n] * in[b, g, k]; row chunks are reduced in a fixed order. */
typedef struct aread_tower_wgrad_args {
  int64_t m;
  int32_t groups;
  int32_t n;                    /* output width of the layer, n * k <= 4096                 */
  int32_t k;                    /* input width                                              */
  const float* dz;              /* fp32 [m, ld_dz]                                          */
  int64_t ld_dz;
  const float* in;              /* fp32 [m, ld_in]                                          */
  int64_t ld_in;
  int64_t in_group_stride;      /* k for side-by-side groups, 0 = shared input              */
  float* d_w;                   /* fp32 [groups, n, k] contiguous                           */
  void* workspace;              /* aread_tower_wgrad_workspace_bytes(m, groups, n, k)       */
  size_t workspace_bytes;
} aread_tower_wgrad_args;

AREAD_API size_t aread_tower_wgrad_workspace_bytes(int64_t m, int32_t groups, int32_t n, int32_t k);
AREAD_API int aread_tower_wgrad(const aread_tower_wgrad_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * HEI gate mixing of one level, row-local (model/aread.py:282-288, and :169-171 without a mask):
 *   s = softmax(logits[b, t, :]);  with a mask: sm = s * edges[t, :], r = sm / (sum sm + 1e-8); else r = s
 *   out[b, t, :] = sum_j r[j] * u_prev[b, slot(j), :]
 * `logits` are the gate Linear outputs for the towers that run (compact, ascending), `prev_slot[j]`
 * is the compact index of previous-level tower j in u_prev (-1: that tower does not run, its output
 * is zero) and `slot_tower` the inverse map.  Forward when d_out == NULL, else backward.
 * `sm` (optional) receives s * edges: its batch mean is the gate value HEMP thresholds (aread.py:290-295).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_gate_mix_args {
  int64_t m;
  int32_t n_tower;          /* towers of this level that run                    */
  int32_t n_prev;           /* all towers of the previous level (<= 32)         */
  int32_t n_prev_active;    /* previous-level towers that run                   */
  int32_t width;            /* output width of the previous level's towers      */
  const float* logits;      /* [m, n_tower, n_prev]                             */
  const float* edges;       /* [n_tower, n_prev] of 0/1, or NULL (no mask)      */
  const int32_t* prev_slot; /* device [n_prev]                                  */
  const int32_t* slot_tower;/* device [n_prev_active] (backward)                */
  const float* u_prev;      /* [m, n_prev_active, width]                        */
  float* out;               /* forward out [m, n_tower, width]                  */
  float* sm;                /* forward optional out [m, n_tower, n_prev]        */
  const float* d_out;       /* backward in [m, n_tower, width]                  */
  float* d_logits;          /* backward out [m, n_tower, n_prev]                */
  float* d_u_prev;          /* backward out [m, n_prev_active, width]           */
  float* r_scratch;         /* unused (kept for layout stability); may be NULL  */
  /* Logits that live inside a wider matrix (the gate Linear evaluated as extra columns of the row pass):         */
  int64_t ld_logits;        /* row stride of `logits` in floats; 0 = n_tower * n_prev (contiguous)               */
  const float* logit_offset;/* optional [n_tower * n_prev] added to the logits (bias + row-independent part)      */
  int64_t ld_dlogits;       /* row stride of `d_logits`; 0 = contiguous                                           */
} aread_gate_mix_args;

AREAD_API int aread_gate_mix(const aread_gate_mix_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Mixed-domain batches (BASELINE.json north_star: "domain-sorted, segment-offset grouped ... skips expert/domain
 * tiles zeroed by the HEMP mask").
 *
 * aread_hei_mixed_eval: model/aread.py:263-322 (hier_tower_mask_forward) + the 'domain_with_mask' head
 * (aread.py:224-234) in EVAL mode for rows of many domains at once: row b runs under the mask of domain[b].  Eval
 * mode is row-local (BatchNorm on running statistics, no dropout), so the result equals the reference called once
 * per domain (run.py:719-727) row for row.  Masks arrive as bit tables per domain:
 *   active[d, l]            bit t = tower t of level l runs        (any edge enters it, aread.py:268)
 *   edges[d, e]             for level l >= 1, tower t: word e = (sum_{1<=l'<l} n_tower[l']) + t, bit j = edge from
 *                           tower j of level l-1 (the mask column mask[l][:, t])
 * Rows whose domain id is outside [0, n_domain) produce y = 0.  (tile, tower) pairs that are pruned for every row
 * of a 32-row tile are skipped, so domain-sorted input skips most pruned work.
 * ---------------------------------------------------------------------------------------------- */
#define AREAD_MIXED_MAX_LEVEL 4
#define AREAD_MIXED_MAX_LAYER 3
typedef struct aread_hei_mixed_args {
  int64_t m;
  int32_t n_level;                 /* <= AREAD_MIXED_MAX_LEVEL                                       */
  int32_t n_layer;                 /* Linear layers per tower MLP, <= AREAD_MIXED_MAX_LAYER          */
  int32_t n_tower[AREAD_MIXED_MAX_LEVEL];
  int32_t width_in;                /* level-0 input width (expert output width)                      */
  int32_t dims[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];   /* tower_dims[l][j], multiples of 4  */
  int32_t n_domain;
  int32_t edge_words;              /* uint32 words per domain in `edges`: sum_{l>=1} n_tower[l]      */
  int32_t bn_skip;                 /* 1: BatchNorm is the identity (batch of one, layer.py:226)      */
  float eps;
  const int32_t* domain;           /* domain id of row b at domain[b * domain_stride]                */
  int64_t domain_stride;
  const uint32_t* active;          /* [n_domain, n_level]                                            */
  const uint32_t* edges;           /* [n_domain, edge_words]                                         */
  const float* t0;                 /* [m, n_tower[0], width_in] level-0 tower inputs (MMoE mixture)  */
  const float* logits[AREAD_MIXED_MAX_LEVEL];      /* l >= 1: gate Linear outputs [m, n_tower[l], n_tower[l-1]] */
  const float* weight[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];        /* packed [n_tower[l], N, K]       */
  const float* bias[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];          /* [n_tower[l], N] each            */
  const float* gamma[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];
  const float* beta[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];
  const float* running_mean[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];
  const float* running_var[AREAD_MIXED_MAX_LEVEL][AREAD_MIXED_MAX_LAYER];
  const float* head_cross;         /* [m, n_last] cross-network part of the heads (aread_rowpass_fwd) */
  const float* lin;                /* [m]                                                            */
  const float* w_tail;             /* [n_last, w_last] columns of towers_linear on the tower output  */
  float* y;                        /* out [m]: mean over the row's active heads                      */
  float* y_stack;                  /* optional out [n_last, m]: per-head probabilities (0: pruned)   */
} aread_hei_mixed_args;

AREAD_API int aread_hei_mixed_eval_supported(const aread_hei_mixed_args* args);
AREAD_API int aread_hei_mixed_eval(const aread_hei_mixed_args* args, aread_stream_t stream);

/* mean[d, :] = mean over the rows b with domain[b] == d of values[b, :], rows added in batch order (deterministic);
 * count[d] = number of such rows (mean = 0 when there are none).  Replaces the per-domain boolean-index loop that
 * records unmasked gate values for a mixed batch (model/aread.py:187-200).                                       */
#define AREAD_DOMAIN_MEAN_MAX_COLS 512
typedef struct aread_domain_mean_args {
  int64_t m;
  int32_t width;                   /* columns, <= AREAD_DOMAIN_MEAN_MAX_COLS */
  int32_t n_domain;
  const float* values;             /* [m, ld]                                */
  int64_t ld;
  const int32_t* domain;           /* domain[b * domain_stride]              */
  int64_t domain_stride;
  float* mean;                     /* out [n_domain, width]                  */
  int32_t* count;                  /* optional out [n_domain]                */
} aread_domain_mean_args;

AREAD_API int aread_domain_mean(const aread_domain_mean_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Adam step over a list of fp32 tensors in one launch (coupled weight decay).  Replaces
 * torch.optim.Adam.step as the trainer configures it (run.py:830-831: betas (0.9, 0.99), eps 1e-8,
 * weight_decay 1e-8) -- SURVEY.md 8(f) rank 1.  Tensors without a gradient this step are left out of
 * the list by the caller (moments and step counts untouched).  Chunking as in aread_l2_reg_args with
 * aread_adam_chunk() elements per chunk.
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_adam_args {
  int32_t n_tensors;
  int64_t n_chunks;
  float* const* params;          /* device arrays [n_tensors] of device pointers        */
  const float* const* grads;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  const int64_t* sizes;          /* device [n_tensors]                                  */
  const int64_t* chunk_start;    /* device [n_tensors]                                  */
  const float* step_size;        /* device [n_tensors]: lr / (1 - beta1^t)              */
  const float* bc2_sqrt;         /* device [n_tensors]: sqrt(1 - beta2^t)               */
  float beta1, beta2, eps, weight_decay;
  const float* l2_twice;         /* device [n_tensors] or NULL: 2 * l2 of an L2 regulariser folded into  */
                                 /* the step (g += l2_twice * p before the weight decay); grads[t] may   */
                                 /* be NULL for such tensors (a gradient of zero)                        */
  /* Device-side step counters (CUDA-graph replay: nothing about the launch changes from step to step).  When
     step_counts != NULL, step_size / bc2_sqrt are ignored: the launch first adds 1 to step_counts[slot[t]] for every
     tensor of the list (its own tiny kernel) and every chunk derives lr / (1 - beta1^t) and sqrt(1 - beta2^t) from
     the counter in double precision.                                                                              */
  float* step_counts;            /* device [any]: one fp32 counter per parameter                         */
  const int64_t* slot;           /* device [n_tensors]: index of tensor t's counter                      */
  double lr, beta1_d, beta2_d;   /* the hyper-parameters as the host holds them (Python floats are doubles): the  */
                                 /* corrections then equal the host-computed ones bit for bit             */
} aread_adam_args;

AREAD_API int64_t aread_adam_chunk(void);
AREAD_API int aread_adam_step(const aread_adam_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-tensor copy / fill in ONE launch: dst[t][0:bytes[t]] = src[t][...]  (src == NULL: zero fill).
 * Serves AREAD.load_model_state (model/aread.py:545-546: ~300 restores of every rolled-back tensor per regroup,
 * run.py:631) and FusedAdam.reset (a fresh `optimizer_fast` per candidate, run.py:632-633, without reallocating
 * its moments) -- SURVEY.md 8(f) rank 2.  Byte counts and pointers must be multiples of 4; chunks are
 * aread_multi_copy_chunk() bytes, chunk_start[t] = first chunk of tensor t (prefix sum of ceil(bytes / chunk)).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_multi_copy_args {
  int32_t n_tensors;
  int64_t n_chunks;
  void* const* dst;              /* device arrays [n_tensors] of device pointers        */
  const void* const* src;        /* device [n_tensors], or NULL: zero fill              */
  const int64_t* bytes;          /* device [n_tensors]                                  */
  const int64_t* chunk_start;    /* device [n_tensors]                                  */
} aread_multi_copy_args;

AREAD_API int64_t aread_multi_copy_chunk(void);
AREAD_API int aread_multi_copy(const aread_multi_copy_args* args, aread_stream_t stream);
/* Same table, but dst[t] receives bf16(src[t]) of the fp32 source (bytes[t] = SOURCE bytes, a multiple of 16;
 * pointers 16-byte aligned): the bf16 operand copies of all expert weights in one launch per step.            */
AREAD_API int aread_multi_cast_bf16(const aread_multi_copy_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Bagging loss of the HEI heads: loss[0] = (1 / n_tower) * sum_t mean_b BCE(probs[t, b], labels[b])
 * and d loss / d probs in the same pass.  Replaces the trainer's per-tower BCELoss sum (run.py:643-644,
 * 672-677, criterion run.py:833); element arithmetic as torch.nn.BCELoss (logs clamped at -100).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_bagging_bce_args {
  int64_t m;             /* samples                                         */
  int32_t n_tower;       /* active last-level towers (rows of probs)        */
  const float* probs;    /* [n_tower, m], in [0, 1]                         */
  const float* labels;   /* [m] fp32                                        */
  float* loss;           /* device scalar                                   */
  float* d_probs;        /* [n_tower, m] or NULL                            */
  void* workspace;       /* aread_bagging_bce_workspace_bytes(m, n_tower)   */
  size_t workspace_bytes;
} aread_bagging_bce_args;

AREAD_API size_t aread_bagging_bce_workspace_bytes(int64_t m, int32_t n_tower);
AREAD_API int aread_bagging_bce(const aread_bagging_bce_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Output heads of the active last-level towers (model/aread.py:307-310):
 *   probs[t, b] = sigmoid(head_cross[b, t] + <h[b, t, :], w_tail[t, :]> + lin[b])
 * head_cross is the cross-network part of towers_linear (aread_rowpass_fwd), w_tail the columns of
 * towers_linear that multiply the tower output.  Forward when d_probs == NULL; else
 *   dz[b, t] = d_probs[t, b] * p (1 - p), d_lin[b] = sum_t dz, d_h[b, t, :] = dz * w_tail[t, :],
 *   d_w_tail[t, :] = sum_b dz[b, t] * h[b, t, :] (fixed reduction order).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aread_head_args {
  int64_t m;
  int32_t n_tower, width;
  const float* head_cross;   /* [m, n_tower]            (forward)              */
  const float* lin;          /* [m]                     (forward)              */
  const float* h;            /* [m, n_tower, width]                            */
  const float* w_tail;       /* [n_tower, width]                               */
  float* probs;              /* [n_tower, m]: out (forward) / in (backward)    */
  const float* d_probs;      /* [n_tower, m]            (backward)             */
  float* dz;                 /* out [m, n_tower]                               */
  float* d_lin;              /* out [m]                                        */
  float* d_h;                /* out [m, n_tower, width]                        */
  float* d_w_tail;           /* optional out [n_tower, width]: sum_b dz[b, t] * h[b, t, :] */
  void* workspace;           /* aread_head_workspace_bytes(n_tower, width) when d_w_tail is set */
  size_t workspace_bytes;
} aread_head_args;

AREAD_API size_t aread_head_workspace_bytes(int32_t n_tower, int32_t width);
AREAD_API int aread_head(const aread_head_args* args, aread_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CUDA IPC plumbing of the row-sharded table (one process per GPU).  aread_ipc_export gives the
 * 64-byte handle of the allocation that contains `ptr` and ptr's offset inside it; a peer process
 * passes both to aread_ipc_open (with ITS compute device) and receives a pointer its kernels can
 * dereference over NVLink -- that pointer goes into aread_gather_args.shards.
 * ---------------------------------------------------------------------------------------------- */
#define AREAD_IPC_HANDLE_BYTES 64
AREAD_API int aread_ipc_export(const void* ptr, unsigned char* handle_out, int64_t* offset_out);
AREAD_API int aread_ipc_open(const unsigned char* handle, int64_t offset, int32_t device, void** ptr_out);
AREAD_API int aread_ipc_close(void* ptr, int64_t offset);

#ifdef __cplusplus
}
#endif
#endif /* AREAD_SM100_H */
