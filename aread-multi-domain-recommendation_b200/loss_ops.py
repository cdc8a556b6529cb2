"""Bagging BCE of the per-tower probabilities as one autograd node on csrc/loss.cu.

`bagging_bce(y_stack, y)` equals the trainer's
    sum(criterion(y_stack[t], y) for t in range(n_act)) / n_act        (run.py:643-644, 672-677)
with criterion = torch.nn.BCELoss() (run.py:833)."""
import ctypes

import torch

from . import _lib
from . import _mem


class BaggingBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, labels):
        if not probs.is_cuda:
            raise RuntimeError("aread_b200: bagging_bce needs CUDA tensors (no CPU fallback)")
        if probs.dim() != 2 or labels.numel() != probs.shape[1]:
            raise ValueError(f"expected probs [n_tower, B] and B labels, got {tuple(probs.shape)} and {tuple(labels.shape)}")
        probs = probs.contiguous().float()
        labels = labels.reshape(-1).to(torch.float32).contiguous()
        T, m = probs.shape
        dev = probs.device
        lib = _lib.load()
        need = int(lib.aread_bagging_bce_workspace_bytes(m, T))
        ws = _mem.workspace("bagging_bce", dev, max(need, 4096))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        want_grad = ctx.needs_input_grad[0]
        d_probs = torch.empty_like(probs) if want_grad else None
        args = _lib.BaggingBceArgs(m, T, probs.data_ptr(), labels.data_ptr(), loss.data_ptr(),
                                   d_probs.data_ptr() if want_grad else None, ws.data_ptr(), ws.numel())
        _lib.check(lib.aread_bagging_bce(ctypes.byref(args), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.d_probs = d_probs
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.d_probs * g if ctx.d_probs is not None else None), None


def bagging_bce(y_stack, targets):
    """y_stack [n_act, B] probabilities (mode='domain_mask_bagging'), targets [B] or [B, 1] (any real dtype)."""
    return BaggingBCE.apply(y_stack, targets)
