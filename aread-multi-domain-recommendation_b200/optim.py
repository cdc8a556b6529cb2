"""FusedAdam: torch.optim.Adam's arithmetic (coupled weight decay, no amsgrad) as ONE kernel launch over
all parameter tensors (csrc/adam.cu).  Same constructor signature and state_dict layout as
torch.optim.Adam, so it can replace `torch.optim.Adam(model.parameters(), lr=..., betas=..., eps=...,
weight_decay=...)` at run.py:830 / 632 without other changes.  Parameters whose `.grad` is None are
skipped entirely (moments, step count and weights untouched), exactly like torch."""
import ctypes
import math

import numpy as np
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._chunk = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if self._chunk is None:
            self._chunk = int(lib.aread_adam_chunk())
        chunk = self._chunk
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            todo = []
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse or not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam handles dense fp32 CUDA parameters only")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)       # host scalar, like torch's default
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if not g.is_contiguous():
                    g = g.contiguous()
                todo.append((p, g, st))
            if not todo:
                continue
            n = len(todo)
            device = todo[0][0].device
            # rows: params, grads, exp_avg, exp_avg_sq, sizes, chunk_start (int64); step_size, bc2_sqrt (fp32).
            # A fresh pinned block per step: torch's host allocator will not recycle it before the copy ran.
            host = torch.empty((8, n), dtype=torch.int64, pin_memory=True)
            dev = torch.empty((8, n), dtype=torch.int64, device=device)
            h = host.numpy()
            f = h.view(np.float32).reshape(8, -1)                               # fp32 view of the same rows
            start = 0
            for i, (p, g, st) in enumerate(todo):
                t = float(st["step"])
                h[0, i], h[1, i] = p.data_ptr(), g.data_ptr()
                h[2, i], h[3, i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                h[4, i], h[5, i] = p.numel(), start
                start += (p.numel() + chunk - 1) // chunk
                f[6, i] = lr / (1.0 - beta1 ** t)
                f[7, i] = math.sqrt(1.0 - beta2 ** t)
            dev.copy_(host, non_blocking=True)
            base, row = dev.data_ptr(), dev.stride(0) * 8
            args = _lib.AdamArgs(n, start, base, base + row, base + 2 * row, base + 3 * row, base + 4 * row,
                                 base + 5 * row, base + 6 * row, base + 7 * row, beta1, beta2, eps, wd)
            _lib.check(lib.aread_adam_step(ctypes.byref(args),
                                           ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
            self._keep = [g for _, g, _ in todo]       # contiguous copies stay alive until the kernel has run
        return loss
