"""L2 regulariser as one autograd node on csrc/reg_loss.cu (reference: model/layer.py:96-112)."""
import ctypes

import torch

from . import _lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Plan:
    """Device-side description of the tensor list, rebuilt only when a data pointer changes."""

    def __init__(self, tensors, l2s, device):
        chunk = int(_lib.load().aread_l2_reg_chunk())
        sizes = [t.numel() for t in tensors]
        starts, n = [], 0
        for s in sizes:
            starts.append(n)
            n += (s + chunk - 1) // chunk
        self.n_chunks = n
        self.key = tuple(t.data_ptr() for t in tensors)
        self.ptrs = torch.tensor(self.key, dtype=torch.int64, device=device)
        self.sizes = torch.tensor(sizes, dtype=torch.int64, device=device)
        self.l2 = torch.tensor(l2s, dtype=torch.float32, device=device)
        self.chunk_start = torch.tensor(starts, dtype=torch.int64, device=device)
        self.workspace = torch.empty(max(n, 1), dtype=torch.float32, device=device)
        self.numel = sizes
        # gradient buffer of the backward, one slice per tensor.  It is reused every step: the plan keeps the
        # views referenced, so autograd never adopts one as a .grad in place (it copies / adds out of place).
        self.flat_grad = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        self.grads, off = [], 0
        for t, s in zip(tensors, sizes):
            self.grads.append(self.flat_grad[off:off + s].view(t.shape))
            off += s
        self.grad_ptrs = torch.tensor([g.data_ptr() for g in self.grads], dtype=torch.int64, device=device)
        self.seq = 0


_PLANS = {}


class L2Reg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, l2s, *tensors):
        dev = tensors[0].device
        key = tuple(t.data_ptr() for t in tensors)
        plan = _PLANS.get(key)
        if plan is None:
            if len(_PLANS) > 8:
                _PLANS.clear()
            plan = _Plan(tensors, l2s, dev)
            _PLANS[key] = plan
        out = torch.empty(1, dtype=torch.float32, device=dev)
        args = _lib.L2RegArgs(len(tensors), plan.n_chunks, plan.ptrs.data_ptr(), None, plan.sizes.data_ptr(),
                              plan.l2.data_ptr(), plan.chunk_start.data_ptr(), out.data_ptr(),
                              plan.workspace.data_ptr(), plan.workspace.numel() * 4)
        _lib.check(_lib.load().aread_l2_reg_fwd(ctypes.byref(args), _stream(dev)))
        ctx.plan = plan
        plan.seq += 1
        ctx.seq = plan.seq
        return out

    @staticmethod
    def backward(ctx, g_out):
        plan = ctx.plan
        dev = g_out.device
        g_out = g_out.contiguous()
        if ctx.seq == plan.seq:          # the latest evaluation owns the plan's buffer
            grads, gptrs = plan.grads, plan.grad_ptrs
        else:                            # an older graph that is still alive: private buffer
            flat = torch.empty_like(plan.flat_grad)
            grads, off = [], 0
            for g in plan.grads:
                grads.append(flat[off:off + g.numel()].view(g.shape))
                off += g.numel()
            gptrs = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64, device=dev)
        args = _lib.L2RegArgs(len(grads), plan.n_chunks, plan.ptrs.data_ptr(), gptrs.data_ptr(),
                              plan.sizes.data_ptr(), plan.l2.data_ptr(), plan.chunk_start.data_ptr(), None, None, 0)
        _lib.check(_lib.load().aread_l2_reg_bwd(ctypes.byref(args), ctypes.c_void_p(g_out.data_ptr()), _stream(dev)))
        return (None, *grads)


def regularization_loss(regularization_weight, device):
    """sum_groups sum_w l2 * sum(w^2) as a [1] tensor.  L1 terms (never registered by AREAD) fall back to
    elementwise torch ops."""
    tensors, l2s, extra = [], [], None
    for weights, l1, l2 in regularization_weight:
        for w in weights:
            p = w[1] if isinstance(w, tuple) else w
            if l2 > 0:
                tensors.append(p)
                l2s.append(float(l2))
            if l1 > 0:
                term = torch.sum(l1 * torch.abs(p)).reshape(1)
                extra = term if extra is None else extra + term
    if not tensors:
        return extra if extra is not None else torch.zeros((1,), device=device)
    if not tensors[0].is_cuda:
        raise RuntimeError("aread_b200: parameters must live on a CUDA device (no CPU fallback)")
    if not all(t.is_contiguous() for t in tensors):
        tensors = [t.contiguous() for t in tensors]
    out = L2Reg.apply(tuple(l2s), *tensors)
    return out if extra is None else out + extra
