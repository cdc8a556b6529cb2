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
                   a_lo=None, w_lo=None):
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
        a_lo.data_ptr() if a_lo is not None else None, w_lo.data_ptr() if w_lo is not None else None)
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


def bn_act_bwd(z, d_out, saved, bn_skip, p, seed, salt, dz_dtype=torch.bfloat16, want_lo=False):
    """(dz, d_gamma, d_beta, d_bias) for out = dropout(relu(bn(z))); with want_lo dz is (hi, lo)."""
    m, width = z.shape
    grads = torch.empty((3, width), dtype=torch.float32, device=z.device)     # parameter gradients: never arena
    dz = _mem.empty((m, width), dz_dtype, z.device)
    dz_lo = _mem.empty((m, width), torch.bfloat16, z.device) if want_lo else None
    ws = _bn_workspace(z.device, width)
    args = _lib.BnActBwdArgs(m, width, 1 if bn_skip else 0, float(p), salt, seed, z.data_ptr(), z.stride(0),
                             d_out.data_ptr(), d_out.stride(0), *_rows(saved, 4), *_rows(grads, 3),
                             _ptr(dz) if dz_dtype == torch.float32 else None,
                             _ptr(dz) if dz_dtype == torch.bfloat16 else None, width, ws.data_ptr(), ws.numel(),
                             _ptr(dz_lo), SEED_PTR)
    _lib.check(_lib.load().aread_bn_act_bwd(ctypes.byref(args), _stream(z.device)))
    return ((dz, dz_lo) if want_lo else dz), grads[0], grads[1], grads[2]


def mmoe_mix_fwd(z, saved, gate, n_expert, n_gate, p, seed, salt):
    m = z.shape[0]
    width = z.shape[1] // n_expert
    out = _mem.empty((m, n_gate, width), torch.float32, z.device)
    rows = _rows(saved, 4)
    args = _lib.MmoeMixArgs(m, width, n_expert, n_gate, float(p), seed, salt, z.data_ptr(), z.stride(0),
                            rows[2], rows[3], gate.data_ptr(), out.data_ptr(), None, None, None, SEED_PTR)
    _lib.check(_lib.load().aread_mmoe_mix(ctypes.byref(args), _stream(z.device)))
    return out


def mmoe_mix_bwd(z, saved, gate, d_out, n_expert, n_gate, p, seed, salt):
    m = z.shape[0]
    width = z.shape[1] // n_expert
    d_h = _mem.empty((m, n_expert * width), torch.float32, z.device)
    d_gate = _mem.empty((m, n_gate, n_expert), torch.float32, z.device)
    rows = _rows(saved, 4)
    args = _lib.MmoeMixArgs(m, width, n_expert, n_gate, float(p), seed, salt, z.data_ptr(), z.stride(0),
                            rows[2], rows[3], gate.data_ptr(), None, d_out.data_ptr(), d_h.data_ptr(),
                            d_gate.data_ptr(), SEED_PTR)
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
    it -- bench.py times `run()` alone for the roofline line."""
    m = a_op.shape[0]
    bias = torch.zeros(groups * n, dtype=torch.float32, device=a_op.device)
    out = torch.empty(m, groups * n, dtype=torch.float32, device=a_op.device)

    def run():
        grouped_linear(a_op, w_op, bias, n, k, groups, 0, out=out)
    return run, m * groups * n * 4, "grouped_linear_kernel<128> (tcgen05 / TMEM, TMA in, TMA out, fp32 output)"
