"""Autograd node for the row pass (csrc/rowpass.cu): linear term, MMoE gate softmax, cross network
and the cross part of the output heads as one skinny fp32 product over the embedding row."""
import ctypes

import torch

from . import _lib
from . import _mem


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _args(m, e, layout, ldp, **ptrs):
    a = _lib.RowpassArgs()
    a.m, a.e = m, e
    a.n_gate, a.n_expert, a.n_cross, a.n_head = layout
    a.ldp = ldp
    for k, v in ptrs.items():
        setattr(a, k, v.data_ptr() if v is not None else None)
    return a


class RowPass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, offset, layout):
        n_gate, n_expert, n_cross, n_head = layout
        m, e = x.shape
        nj = 1 + n_gate * n_expert + n_cross + n_head
        assert w.shape == (nj, e) and offset.shape == (nj,)
        ldp = (nj + 3) // 4 * 4
        dev = x.device
        x, w, offset = x.contiguous(), w.contiguous(), offset.contiguous()
        p = torch.empty((m, ldp), dtype=torch.float32, device=dev)
        lin = torch.empty((m,), dtype=torch.float32, device=dev)
        gate = torch.empty((m, n_gate, n_expert), dtype=torch.float32, device=dev)
        alpha = torch.empty((m, n_cross + 1), dtype=torch.float32, device=dev)
        head = torch.empty((m, n_head), dtype=torch.float32, device=dev)
        a = _args(m, e, layout, ldp, x=x, w=w, offset=offset, p=p, lin=lin, gate=gate, alpha=alpha, head=head)
        _lib.check(_lib.load().aread_rowpass_fwd(ctypes.byref(a), _stream(dev)))
        ctx.layout, ctx.ldp = layout, ldp
        ctx.save_for_backward(x, w, p, gate, alpha)
        ctx.mark_non_differentiable(alpha)
        return lin, gate, head, alpha

    @staticmethod
    def backward(ctx, d_lin, d_gate, d_head, _d_alpha):
        x, w, p, gate, alpha = ctx.saved_tensors
        m, e = x.shape
        nj = w.shape[0]
        dev = x.device
        d_p = torch.empty((m, ctx.ldp), dtype=torch.float32, device=dev)
        d_c = torch.empty((m, ctx.ldp), dtype=torch.float32, device=dev)
        d_x = torch.empty((m, e), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        d_w = torch.empty((nj, e), dtype=torch.float32, device=dev)
        need = int(_lib.load().aread_rowpass_workspace_bytes(m, e, nj))
        ws = _mem.workspace("rowpass", dev, need)
        a = _args(m, e, ctx.layout, ctx.ldp, x=x, w=w, p=p, gate=gate, alpha=alpha,
                  d_lin=d_lin.contiguous() if d_lin is not None else None,
                  d_gate=d_gate.contiguous() if d_gate is not None else None,
                  d_head=d_head.contiguous() if d_head is not None else None,
                  d_p=d_p, d_c=d_c, d_x=d_x, d_w=d_w, workspace=ws)
        a.workspace_bytes = ws.numel()
        _lib.check(_lib.load().aread_rowpass_bwd(ctypes.byref(a), _stream(dev)))
        d_offset = d_c[:, :nj].sum(dim=0)
        return d_x, d_w, d_offset, None


def rowpass(model, X, active0, active_last):
    """(lin [B], gate [B, n_active0, n_expert], head_cross [B, n_active_last]) for the active level-0
    towers / last-level towers given by index lists.  The stacked weight / offset vectors are built
    with torch.cat from the module parameters so that autograd routes the gradients back (and leaves
    the parameters of inactive towers without gradient, like the reference)."""
    E = X.shape[1]
    n_expert = len(model.mmoe_experts)
    n_cross = model.cn.num_layers
    w_rows = [model.linear.fc.weight]
    w_rows += [model.mmoe_gates[g][0].weight for g in active0]
    w_rows += [lin.weight for lin in model.cn.w]
    w_out = [model.towers_linear[t].weight[:, :E] for t in active_last]
    w_rows += w_out
    w = torch.cat(w_rows, dim=0)
    # beta_k = b_0 + .. + b_{k-1} (beta_0 = 0): the row-independent part of the k-th cross state
    beta, off = None, [model.linear.fc.bias]
    off += [model.mmoe_gates[g][0].bias for g in active0]
    kappas = []
    for k in range(n_cross):
        kappas.append((model.cn.w[k].weight[0] * beta).sum().reshape(1) if beta is not None
                      else torch.zeros(1, dtype=X.dtype, device=X.device))
        beta = model.cn.b[k] if beta is None else beta + model.cn.b[k]
    off += kappas
    if w_out:
        off.append(torch.cat(w_out, dim=0) @ beta if beta is not None
                   else torch.zeros(len(w_out), dtype=X.dtype, device=X.device))
    offset = torch.cat(off, dim=0)
    lin, gate, head, _ = RowPass.apply(X, w, offset, (len(active0), n_expert, n_cross, len(active_last)))
    return lin, gate, head
