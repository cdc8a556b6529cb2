"""Thin torch-tensor wrappers over the dense C-ABI entry points (no arithmetic here)."""
import ctypes

import torch

from . import _lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def grouped_linear(a, w, bias, n, k, groups, a_group_cols=0, group_mask=None, out=None, out_dtype=torch.float32):
    """C[:, g*n:(g+1)*n] = A[:, g*a_group_cols : +k] @ W[g*n:(g+1)*n, :k]^T (+ bias) for active groups.

    a: bf16 [m, lda]; w: bf16 [groups*n, ldb]; bias: fp32 [groups*n] or None.  Columns of masked-out
    groups are left untouched in `out` (zero-initialised when `out` is None)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_cuda and w.is_cuda
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m = a.shape[0]
    full = (1 << groups) - 1
    mask = full if group_mask is None else int(group_mask) & full
    if out is None:
        alloc = torch.empty if mask == full else torch.zeros
        out = alloc((m, groups * n), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and out.dtype in (torch.float32, torch.bfloat16)
    args = _lib.GroupedLinearArgs(
        m, n, k, groups, a_group_cols, mask, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0),
        bias.data_ptr() if bias is not None else None,
        out.data_ptr() if out.dtype == torch.float32 else None,
        out.data_ptr() if out.dtype == torch.bfloat16 else None, out.stride(0))
    _lib.check(_lib.load().aread_grouped_linear_bf16(ctypes.byref(args), _stream(a.device)))
    return out


_WGRAD_WS = {}


def grouped_wgrad(dz, a, n, k, groups, a_group_cols=0, group_mask=None, out=None):
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
                                 a.stride(0), out.data_ptr(), None, 0)
    need = int(_lib.load().aread_grouped_wgrad_workspace_bytes(ctypes.byref(args)))
    ws = _WGRAD_WS.get(dz.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dz.device)
        _WGRAD_WS[dz.device] = ws
    args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(_lib.load().aread_grouped_wgrad_bf16(ctypes.byref(args), _stream(dz.device)))
    return out
