"""Thin torch-tensor wrappers over the dense C-ABI entry points (no arithmetic here)."""
import ctypes

import torch

from . import _lib
from . import _mem


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def split_bf16(x):
    """x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the operands of the split-precision GEMMs."""
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def grouped_linear(a, w, bias, n, k, groups, a_group_cols=0, group_mask=None, out=None, out_dtype=torch.float32,
                   a_lo=None, w_lo=None, lo_lo=False):
    """C[:, g*n:(g+1)*n] = A[:, g*a_group_cols : +k] @ W[g*n:(g+1)*n, :k]^T (+ bias) for active groups.

    a: bf16 [m, lda]; w: bf16 [groups*n, ldb]; bias: fp32 [groups*n] or None.  Columns of masked-out
    groups are left untouched in `out` (zero-initialised when `out` is None)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_cuda and w.is_cuda
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m = a.shape[0]
    full = (1 << groups) - 1
    mask = full if group_mask is None else int(group_mask) & full
    if out is None:
        alloc = _mem.empty if mask == full else _mem.zeros
        out = alloc((m, groups * n), out_dtype, a.device)
    assert out.stride(1) == 1 and out.dtype in (torch.float32, torch.bfloat16)
    args = _lib.GroupedLinearArgs(
        m, n, k, groups, a_group_cols, mask, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
        bias.data_ptr() if bias is not None else None,
        out.data_ptr() if out.dtype == torch.float32 else None,
        out.data_ptr() if out.dtype == torch.bfloat16 else None, out.stride(0),
        a_lo.data_ptr() if a_lo is not None else None, w_lo.data_ptr() if w_lo is not None else None,
        1 if (lo_lo and a_lo is not None) else 0)
    if a_lo is not None:
        assert a_lo.stride() == a.stride() and w_lo.stride() == w.stride()
    _lib.check(_lib.load().aread_grouped_linear_bf16(ctypes.byref(args), _stream(a.device)))
    return out



def grouped_wgrad(dz, a, n, k, groups, a_group_cols=0, group_mask=None, out=None, dz_lo=None, a_lo=None):
    """dW[g*n + j, i] = sum_b dz[b, g*n + j] * a[b, g*a_group_cols + i] (fp32 [groups*n, k]).
    Rows of masked-out groups are left untouched (zero when `out` is None)."""
    assert dz.dtype == torch.bfloat16 and a.dtype == torch.bfloat16 and dz.is_cuda and a.is_cuda
    assert dz.stride(1) == 1 and a.stride(1) == 1 and dz.shape[0] == a.shape[0]
    m = dz.shape[0]
    full = (1 << groups) - 1
    mask = full if group_mask is None else int(group_mask) & full
    if out is None:
        alloc = torch.empty if mask == full else torch.zeros
        out = alloc((groups * n, k), dtype=torch.float32, device=dz.device)
    assert out.is_contiguous() and out.dtype == torch.float32
    args = _lib.GroupedWgradArgs(m, n, k, groups, a_group_cols, mask, dz.data_ptr(), dz.stride(0), a.data_ptr(),
                                 a.stride(0), out.data_ptr(), None, 0,
                                 dz_lo.data_ptr() if dz_lo is not None else None,
                                 a_lo.data_ptr() if a_lo is not None else None)
    need = int(_lib.load().aread_grouped_wgrad_workspace_bytes(ctypes.byref(args)))
    ws = _mem.workspace("wgrad", dz.device, need)
    args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(_lib.load().aread_grouped_wgrad_bf16(ctypes.byref(args), _stream(dz.device)))
    return out


# device address of the dropout seed while the fused node records / replays a CUDA graph (fused.py); None: the
# seed travels by value in the launch arguments
SEED_PTR = None

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _bn_workspace(device, width):
    need = int(_lib.load().aread_bn_workspace_bytes(width))
    return _mem.workspace("bn", device, need)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _rows(t, n):
    """Device addresses of the first n rows of a contiguous 2-D fp32 tensor (no view objects)."""
    base, step = t.data_ptr(), t.stride(0) * 4
    return [base + i * step for i in range(n)]


def bn_act_fwd(z, gamma, beta, running_mean, running_var, training, bn_skip, p, seed, salt, out_dtype,
               want_lo=False):
    """(out, saved) or (out, out_lo, saved) with want_lo: out = dropout(relu(bn(z))) as `out_dtype`
    (None: statistics only); saved = [4, width] fp32 rows (mean, rstd, scale, shift)."""
    m, width = z.shape
    saved = _mem.empty((4, width), torch.float32, z.device)
    out = _mem.empty((m, width), out_dtype, z.device) if out_dtype is not None else None
    out_lo = _mem.empty((m, width), torch.bfloat16, z.device) if want_lo else None
    ws = _bn_workspace(z.device, width)
    args = _lib.BnActArgs(m, width, 1 if training else 0, 1 if bn_skip else 0, BN_MOMENTUM, BN_EPS,
                          float(p) if training else 0.0, seed, salt, z.data_ptr(), z.stride(0), _ptr(gamma),
                          _ptr(beta), _ptr(running_mean), _ptr(running_var), *_rows(saved, 4),
                          _ptr(out) if out_dtype == torch.float32 else None,
                          _ptr(out) if out_dtype == torch.bfloat16 else None, width, ws.data_ptr(), ws.numel(),
                          _ptr(out_lo), SEED_PTR)
    _lib.check(_lib.load().aread_bn_act_fwd(ctypes.byref(args), _stream(z.device)))
    return (out, out_lo, saved) if want_lo else (out, saved)


def bn_act_bwd(z, d_out, saved, bn_skip, p, seed, salt, dz_dtype=torch.bfloat16, want_lo=False, dz_out=None):
    """(dz, d_gamma, d_beta, d_bias) for out = dropout(relu(bn(z))); with want_lo dz is (hi, lo).  `dz_out`: a
    preallocated [m, width] view (any row stride) that receives dz."""
    m, width = z.shape
    grads = torch.empty((3, width), dtype=torch.float32, device=z.device)     # parameter gradients: never arena
    dz = dz_out if dz_out is not None else _mem.empty((m, width), dz_dtype, z.device)
    dz_lo = _mem.empty((m, width), torch.bfloat16, z.device) if want_lo else None
    ws = _bn_workspace(z.device, width)
    z16 = z.dtype == torch.bfloat16
    args = _lib.BnActBwdArgs(m, width, 1 if bn_skip else 0, float(p), salt, seed, None if z16 else z.data_ptr(),
                             z.stride(0), d_out.data_ptr(), d_out.stride(0), *_rows(saved, 4), *_rows(grads, 3),
                             _ptr(dz) if dz_dtype == torch.float32 else None,
                             _ptr(dz) if dz_dtype == torch.bfloat16 else None, dz.stride(0), ws.data_ptr(), ws.numel(),
                             _ptr(dz_lo), SEED_PTR, z.data_ptr() if z16 else None)
    _lib.check(_lib.load().aread_bn_act_bwd(ctypes.byref(args), _stream(z.device)))
    return ((dz, dz_lo) if want_lo else dz), grads[0], grads[1], grads[2]


def mmoe_mix_fwd(z, saved, gate, n_expert, n_gate, p, seed, salt):
    m = z.shape[0]
    width = z.shape[1] // n_expert
    out = _mem.empty((m, n_gate, width), torch.float32, z.device)
    rows = _rows(saved, 4)
    z16 = z.dtype == torch.bfloat16
    args = _lib.MmoeMixArgs(m, width, n_expert, n_gate, float(p), seed, salt, None if z16 else z.data_ptr(), z.stride(0),
                            rows[2], rows[3], gate.data_ptr(), out.data_ptr(), None, None, None, SEED_PTR,
                            z.data_ptr() if z16 else None)
    _lib.check(_lib.load().aread_mmoe_mix(ctypes.byref(args), _stream(z.device)))
    return out


def mmoe_mix_bwd(z, saved, gate, d_out, n_expert, n_gate, p, seed, salt):
    m = z.shape[0]
    width = z.shape[1] // n_expert
    d_h = _mem.empty((m, n_expert * width), torch.float32, z.device)
    d_gate = _mem.empty((m, n_gate, n_expert), torch.float32, z.device)
    rows = _rows(saved, 4)
    z16 = z.dtype == torch.bfloat16
    args = _lib.MmoeMixArgs(m, width, n_expert, n_gate, float(p), seed, salt, None if z16 else z.data_ptr(), z.stride(0),
                            rows[2], rows[3], gate.data_ptr(), None, d_out.data_ptr(), d_h.data_ptr(),
                            d_gate.data_ptr(), SEED_PTR, z.data_ptr() if z16 else None)
    _lib.check(_lib.load().aread_mmoe_mix(ctypes.byref(args), _stream(z.device)))
    return d_h, d_gate


def dropout_mask(seed, salt, shape, p, device):
    n = 1
    for s in shape:
        n *= s
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(_lib.load().aread_dropout_mask(seed, salt, n, float(p), out.data_ptr(), _stream(device)))
    return out.view(*shape).bool()


def bench_expert_layer1(a_op, w_op, n, k, groups):
    """(run, output bytes per launch, description) of the expert layer-1 GEMM exactly as the fused node launches
    it in training (bf16 experts) -- bench.py times `run()` alone for the roofline line."""
    m = a_op.shape[0]
    z = torch.empty(m, groups * n, dtype=torch.bfloat16, device=a_op.device)
    n_part = int(_lib.load().aread_expert_gemm_partials(m))
    partial = torch.empty(n_part, 2, groups * n, dtype=torch.float32, device=a_op.device)

    def run():
        _expert_gemm(a_op, w_op, n, k, groups, 0, EPI_STATS, out=z, partial=partial)
    return run, m * groups * n * 2 + partial.numel() * 4, \
        "grouped_linear_kernel<128, STATS> (tcgen05 / TMEM, TMA in, bf16 TMA out, BatchNorm column sums in the epilogue)"


# ------------------------------------------------------------------------------------------------------------------
# bf16 experts: BatchNorm bookkeeping inside the GEMM epilogues (csrc/gemm_tc.cu aread_expert_gemm).  The
# pre-activation is stored once, as bf16 and without the bias (BatchNorm removes it; the finalize kernel folds it
# into the running mean / the eval-mode shift).
# ------------------------------------------------------------------------------------------------------------------
EPI_PLAIN, EPI_STATS, EPI_ACT, EPI_BN_BWD, EPI_BF16 = 0, 1, 2, 3, 4


def _expert_gemm(a, w, n, k, groups, a_group_cols, epilogue, *, k_by_n=False, out, partial=None, saved=None, z=None,
                 p=0.0, salt=0, seed=0, bias=None):
    sv = _rows(saved, 4) if saved is not None else (None,) * 4
    args = _lib.ExpertGemmArgs(
        a.shape[0], n, k, groups, a_group_cols, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
        1 if k_by_n else 0, epilogue, _ptr(bias), out.data_ptr() if out.dtype == torch.float32 else None,
        out.data_ptr() if out.dtype == torch.bfloat16 else None, out.stride(0), _ptr(partial), sv[2], sv[3], sv[0], sv[1],
        _ptr(z), z.stride(0) if z is not None else 0, float(p), salt, seed, SEED_PTR)
    _lib.check(_lib.load().aread_expert_gemm(ctypes.byref(args), _stream(a.device)))
    return out


def _partials(m, width, device):
    n_part = int(_lib.load().aread_expert_gemm_partials(m))
    return _mem.empty((n_part, 2, width), torch.float32, device), n_part


def expert_linear_stats(a, w, n, k, groups, a_group_cols):
    """(z16 [m, groups*n] = bf16(A . W^T), partial [tiles, 2, groups*n] column sums / sums of squares)."""
    m = a.shape[0]
    z = _mem.empty((m, groups * n), torch.bfloat16, a.device)
    partial, _ = _partials(m, groups * n, a.device)
    _expert_gemm(a, w, n, k, groups, a_group_cols, EPI_STATS, out=z, partial=partial)
    return z, partial


def expert_linear_act(a, w, n, k, groups, a_group_cols, saved):
    """bf16(relu(A . W^T * scale + shift)): inference, BatchNorm folded into the epilogue."""
    out = _mem.empty((a.shape[0], groups * n), torch.bfloat16, a.device)
    return _expert_gemm(a, w, n, k, groups, a_group_cols, EPI_ACT, out=out, saved=saved)


def expert_bn_finalize(partial, m, width, bias, gamma, beta, running_mean, running_var, training, bn_skip):
    """saved [4, width] = (mean of the bias-free accumulator, rstd, scale, shift); updates the running statistics."""
    dev = (partial if partial is not None else running_mean).device
    saved = _mem.empty((4, width), torch.float32, dev)
    args = _lib.ExpertBnFinalizeArgs(m, width, partial.shape[0] if partial is not None else 0, 1 if training else 0,
                                     1 if bn_skip else 0, BN_MOMENTUM, BN_EPS, _ptr(partial), _ptr(bias), _ptr(gamma),
                                     _ptr(beta), _ptr(running_mean), _ptr(running_var), *_rows(saved, 4))
    _lib.check(_lib.load().aread_expert_bn_finalize(ctypes.byref(args), _stream(dev)))
    return saved


def identity_saved(width, device):
    """saved rows of an already activated tensor: scale 1, shift 0 (relu is idempotent)."""
    saved = _mem.empty((4, width), torch.float32, device)
    args = _lib.ExpertBnFinalizeArgs(1, width, 0, 0, 1, BN_MOMENTUM, BN_EPS, None, None, None, None, None, None,
                                     *_rows(saved, 4))
    _lib.check(_lib.load().aread_expert_bn_finalize(ctypes.byref(args), _stream(device)))
    return saved


def bn16_fwd(z, saved, training, p, seed, salt, want_bits=False):
    """bf16(dropout(relu(z * scale + shift))) from the bf16 pre-activation; with want_bits also the [m, width / 8]
    uint8 map of the elements that passed ReLU and dropout (the backward reads it instead of hashing again)."""
    m, width = z.shape
    out = _mem.empty((m, width), torch.bfloat16, z.device)
    bits = _mem.empty((m, width // 8), torch.uint8, z.device) if want_bits else None
    sv = _rows(saved, 4)
    args = _lib.Bn16Args(m, width, 0, z.data_ptr(), z.stride(0), sv[2], sv[3], float(p) if training else 0.0, salt, seed,
                         SEED_PTR, out.data_ptr(), width, None, 0, None, None, None, 0, _ptr(bits), 1.0)
    _lib.check(_lib.load().aread_bn16(ctypes.byref(args), _stream(z.device)))
    return (out, bits) if want_bits else out


def bn16_bwd(z, dy, saved, coef, bn_skip, out=None, raw=False, p=0.0, salt=0, seed=0, bits=None):
    """dz16 = bf16(scale * (dy - coef0 - xhat * coef1)) from the bf16 pre-activation and the bf16 gradient.  raw=False:
    `dy` already carries the ReLU / dropout mask (BN_BWD epilogue); raw=True: it is the gradient w.r.t. the activated
    output and the mask comes from `bits` (bn16_fwd) or is rebuilt from (z, saved, p, salt, seed).  `out`: a
    preallocated [m, width] bf16 view (any 16-byte aligned row stride)."""
    m, width = z.shape
    if out is None:
        out = _mem.empty((m, width), torch.bfloat16, z.device)
    sv = _rows(saved, 4)
    args = _lib.Bn16Args(m, width, 1 if bn_skip else 0, z.data_ptr(), z.stride(0), sv[2], sv[3], float(p), salt, seed,
                         SEED_PTR, out.data_ptr(), out.stride(0), dy.data_ptr(), dy.stride(0), sv[0], sv[1],
                         coef.data_ptr(), 1 if raw else 0, _ptr(bits), 1.0)
    _lib.check(_lib.load().aread_bn16(ctypes.byref(args), _stream(z.device)))
    return out


def bn16_bwd_stats(z, d_h, saved, bn_skip, p, salt, seed, bits=None):
    """partial [ctas, 2, width]: per-CTA sums of dy and dy * xhat, dy = d_h masked by ReLU / dropout (from `bits`, or
    rebuilt here)."""
    m, width = z.shape
    n_part = int(_lib.load().aread_bn16_partials(m, width))
    partial = _mem.empty((n_part, 2, width), torch.float32, z.device)
    sv = _rows(saved, 4)
    args = _lib.Bn16Args(m, width, 1 if bn_skip else 0, z.data_ptr(), z.stride(0), sv[2], sv[3], float(p), salt, seed,
                         SEED_PTR, None, 0, d_h.data_ptr(), d_h.stride(0), sv[0], sv[1], None, 1, _ptr(bits), 1.0)
    _lib.check(_lib.load().aread_bn16_bwd_stats(ctypes.byref(args), ctypes.c_void_p(partial.data_ptr()), _stream(z.device)))
    return partial


def expert_dgrad_bf16(dz, w, n_out, k, groups):
    """Data gradient of a grouped Linear as bf16: d_h16 [m, groups*n_out] = dZ . W (W read in place, k-by-n operand)."""
    d_h = _mem.empty((dz.shape[0], groups * n_out), torch.bfloat16, dz.device)
    return _expert_gemm(dz, w, n_out, k, groups, k, EPI_BF16 if n_out % 64 == 0 else EPI_PLAIN, k_by_n=True, out=d_h)


def expert_dgrad_bn_bwd(dz, w, n_out, k, groups, z_prev, saved_prev, p, salt_prev, seed):
    """Data gradient of a grouped Linear, dA = dZ . W (W read in place as a k-by-n operand), with the BatchNorm
    backward bookkeeping of the layer BELOW in the epilogue -> (dy16 [m, groups*n_out], partial)."""
    m = dz.shape[0]
    dy = _mem.empty((m, groups * n_out), torch.bfloat16, dz.device)
    partial, _ = _partials(m, groups * n_out, dz.device)
    _expert_gemm(dz, w, n_out, k, groups, k, EPI_BN_BWD, k_by_n=True, out=dy, partial=partial, saved=saved_prev, z=z_prev,
                 p=p, salt=salt_prev, seed=seed)
    return dy, partial


def expert_bn_bwd_finalize(partial, m, width, bn_skip):
    """(coef [2, width], grads [3, width] = d_gamma, d_beta, d_bias) from the partials of a BN_BWD launch."""
    dev = partial.device
    coef = _mem.empty((2, width), torch.float32, dev)
    grads = torch.empty((3, width), dtype=torch.float32, device=dev)               # parameter gradients: never arena
    g = _rows(grads, 3)
    args = _lib.ExpertBnBwdFinalizeArgs(m, width, partial.shape[0], 1 if bn_skip else 0, partial.data_ptr(), g[0], g[1],
                                        g[2], coef.data_ptr())
    _lib.check(_lib.load().aread_expert_bn_bwd_finalize(ctypes.byref(args), _stream(dev)))
    return coef, grads


def expert_dgrad_plain(dz, w, n_out, k):
    """d_x fp32 [m, n_out] = dZ [m, k] . W [k, n_out] (first expert layer: all experts share the input)."""
    out = _mem.empty((dz.shape[0], n_out), torch.float32, dz.device)
    return _expert_gemm(dz, w, n_out, k, 1, 0, EPI_PLAIN, k_by_n=True, out=out)


class WeightCast:
    """bf16 operand copies of the expert weights, refreshed by ONE launch per forward (csrc/adam.cu
    multi_cast_bf16_kernel) instead of one conversion (and, for the data gradient, one transposed copy) per layer."""

    def __init__(self, flats):
        import numpy as np
        lib = _lib.load()
        chunk = int(lib.aread_multi_copy_chunk())
        dev = flats[0].device
        self.src_ptrs = [f.data_ptr() for f in flats]
        self.out = [torch.empty(f.shape, dtype=torch.bfloat16, device=dev) for f in flats]
        table = np.zeros((4, len(flats)), dtype=np.int64)
        start = 0
        for i, (f, o) in enumerate(zip(flats, self.out)):
            nbytes = f.numel() * 4
            if nbytes % 16 or f.data_ptr() % 16 or not f.is_contiguous():
                raise RuntimeError("WeightCast needs contiguous, 16-byte aligned fp32 tensors")
            table[:, i] = (o.data_ptr(), f.data_ptr(), nbytes, start)
            start += (nbytes + chunk - 1) // chunk
        self.n, self.n_chunks, self.device = len(flats), start, dev
        self.table = torch.from_numpy(table).to(dev)

    def run(self):
        base, row = self.table.data_ptr(), self.table.stride(0) * 8
        args = _lib.MultiCopyArgs(self.n, self.n_chunks, base, base + row, base + 2 * row, base + 3 * row)
        _lib.check(_lib.load().aread_multi_cast_bf16(ctypes.byref(args), _stream(self.device)))
        return self.out
