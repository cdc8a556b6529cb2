"""Host side of the embedding lookup: builds the device-resident lookup plan and wraps the
C-ABI gather / scatter in a torch.autograd.Function.  PyTorch is used for memory and streams
only; the arithmetic is in csrc/embedding.cu.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import _mem


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"aread_b200: {what} must live on a CUDA device -- this implementation is "
                           "sm_100a-only and has no CPU fallback")


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class LookupPlan:
    """Device copy of the lookup layout (include/aread_sm100.h: aread_embed_plan) plus the
    scratch buffers the gradient needs.  One per (embedding module, device)."""

    def __init__(self, offsets, multi_hot_flag, seq_maxlen, method, embed_dim, n_rows, device):
        flag = np.asarray(multi_hot_flag, dtype=bool)
        n_cols = len(offsets)
        if flag.size != n_cols:
            # the reference builds offsets positionally (one-hot columns first); a flag vector of a
            # different length cannot describe x
            raise ValueError(f"multi_hot_flag has {flag.size} entries for {n_cols} columns")
        pooled = bool(flag.any()) and method in ("mean", "sum")
        one_hot_cols = np.nonzero(~flag)[0] if pooled else np.arange(n_cols)
        fields = [[int(c)] for c in one_hot_cols]
        div = [1.0] * len(fields)
        if pooled:
            mh_cols = np.nonzero(flag)[0]
            for j in range(len(mh_cols) // seq_maxlen):
                fields.append([int(c) for c in mh_cols[j * seq_maxlen:(j + 1) * seq_maxlen]])
                div.append(float(seq_maxlen) if method == "mean" else 1.0)
        max_src = max(len(f) for f in fields)
        src = np.zeros((len(fields), max_src), dtype=np.int32)
        for i, f in enumerate(fields):
            src[i, :len(f)] = f
        off = np.asarray(offsets, dtype=np.int64)
        if off.max(initial=0) >= 2 ** 31:
            raise ValueError("table offsets exceed int32")
        self.device = device
        self.n_cols, self.n_fields, self.max_src = n_cols, len(fields), int(max_src)
        self.embed_dim, self.n_rows = int(embed_dim), int(n_rows)
        self.col_offset = torch.from_numpy(off.astype(np.int32)).to(device)
        self.field_src = torch.from_numpy(src.reshape(-1)).to(device)
        self.field_nsrc = torch.tensor([len(f) for f in fields], dtype=torch.int32, device=device)
        self.field_div = torch.tensor(div, dtype=torch.float32, device=device)
        self.status = torch.zeros(2, dtype=torch.int32, device=device)
        self._status_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._status_event = None
        self.shards = None               # sharding.TableShards when the table is row-sharded over GPUs

    def c_plan(self):
        return _lib.EmbedPlan(self.n_cols, self.n_fields, self.max_src, self.embed_dim, self.n_rows,
                              self.col_offset.data_ptr(), self.field_src.data_ptr(),
                              self.field_nsrc.data_ptr(), self.field_div.data_ptr())

    def workspace(self, n_lookups):
        need = int(_lib.load().aread_scatter_workspace_bytes(n_lookups, self.embed_dim))
        if need == 0:
            _lib.check(_lib.AREAD_ERR_CUDA)
        return _mem.workspace("scatter", self.device, need)

    # --- bounds reporting -------------------------------------------------------------------
    # torch raises IndexError for idx >= n_rows; the kernel records it in `status`.  'sync' checks
    # after every lookup (one stream sync), 'deferred' checks the previous lookup's flag without
    # blocking, so the error surfaces one call late.
    def post_lookup(self, mode):
        if mode == "off":
            return
        if mode == "sync":
            st = self.status.tolist()
            if st[0]:
                self.status.zero_()
                raise IndexError(f"index out of range in self (row {st[1]} not in [0, {self.n_rows}))")
            return
        if self._status_event is not None and self._status_event.query():
            if int(self._status_host[0]):
                row = int(self._status_host[1])
                self._status_host.zero_()
                self.status.zero_()
                self._status_event = None
                raise IndexError(f"index out of range in self (row {row} not in [0, {self.n_rows}))")
            self._status_event = None
        if self._status_event is None:
            self._status_host.copy_(self.status, non_blocking=True)
            self._status_event = torch.cuda.Event()
            self._status_event.record(torch.cuda.current_stream(self.device))


def gather(plan, table, x, want_bf16=False, want_lo=False, fence=True):
    """[B, n_cols] int32 ids -> [B, n_fields, D] fp32 (and optionally the bf16 copy; with want_lo the
    bf16 result is the (hi, lo) split pair)."""
    B = x.shape[0]
    out = _mem.empty((B, plan.n_fields, plan.embed_dim), torch.float32, x.device)
    shape16 = (B, plan.n_fields * plan.embed_dim)
    out_bf16 = _mem.empty(shape16, torch.bfloat16, x.device) if want_bf16 else None
    out_lo = _mem.empty(shape16, torch.bfloat16, x.device) if (want_bf16 and want_lo) else None
    sh = plan.shards
    if sh is not None and fence:         # fence=False: the caller has already ordered this lookup after the peers' updates
        sh.fence()
    args = _lib.GatherArgs(plan.c_plan(), B, x.data_ptr(), table.data_ptr(), out.data_ptr(),
                           out_bf16.data_ptr() if want_bf16 else None,
                           out_lo.data_ptr() if out_lo is not None else None, plan.status.data_ptr(),
                           sh.shift if sh is not None else 0, sh.ptrs.data_ptr() if sh is not None else None)
    _lib.check(_lib.load().aread_gather_fwd(ctypes.byref(args), _stream_ptr(x.device)))
    return out, ((out_bf16, out_lo) if want_lo else out_bf16)


def scatter(plan, x, d_out, d_table=None, zero_fill=True, want_sorted=False, reduce=True):
    """Deterministic gradient of `gather` w.r.t. the table.  Returns the dense [n_rows, D] gradient
    (and, for the bookkeeping tests, the sorted rows / positions)."""
    B = x.shape[0]
    n = B * plan.n_cols
    sh = plan.shards
    sparse = sh is not None and sh.sparse
    if sparse:                           # reduced rows go straight into the owners' receive slots (sharding.py)
        d_table = None
    elif sh is not None:                 # owner-major full-size buffer, then reduce-scatter to the shard owner
        d_table = sh.grad_buffer(x.device)
    elif d_table is None:
        d_table = torch.empty((plan.n_rows, plan.embed_dim), dtype=torch.float32, device=x.device)
    ws = plan.workspace(n)
    rows = torch.empty(n, dtype=torch.int32, device=x.device) if want_sorted else None
    pos = torch.empty(n, dtype=torch.int32, device=x.device) if want_sorted else None
    args = _lib.ScatterArgs(plan.c_plan(), B, x.data_ptr(), d_out.data_ptr(),
                            d_table.data_ptr() if d_table is not None else None,
                            1 if zero_fill else 0, ws.data_ptr(), ws.numel(),
                            rows.data_ptr() if want_sorted else None, pos.data_ptr() if want_sorted else None,
                            sh.shift if sh is not None else 0, sh.rows if sh is not None else 0,
                            sh.push_ptrs.data_ptr() if sparse else None)
    _lib.check(_lib.load().aread_scatter_bwd(ctypes.byref(args), _stream_ptr(x.device)))
    if sh is not None and reduce:        # reduce=False: the caller runs the reduce-scatter itself (fused.py)
        d_table = sh.reduce_grad(d_table)
    if want_sorted:
        return d_table, rows, pos
    return d_table


class EmbeddingLookup(torch.autograd.Function):
    """FeaturesEmbedding.forward / backward (model/layer.py:160-183) on the C ABI.  Optionally also
    returns the bf16 copy of the flattened output (the experts' GEMM operand), written by the same
    kernel; it carries no gradient."""

    @staticmethod
    def forward(ctx, table, x, plan, want_bf16=False, want_lo=False):
        out, out_bf16 = gather(plan, table, x, want_bf16, want_lo)
        ctx.plan = plan
        ctx.save_for_backward(x)
        if want_bf16 and want_lo:
            ctx.mark_non_differentiable(*out_bf16)
            return (out, *out_bf16)
        if want_bf16:
            ctx.mark_non_differentiable(out_bf16)
            return out, out_bf16
        return out

    @staticmethod
    def backward(ctx, d_out, *unused):
        (x,) = ctx.saved_tensors
        return scatter(ctx.plan, x, d_out.contiguous()), None, None, None, None


def prepare_ids(x, table):
    _require_cuda(table, "the embedding table")
    _require_cuda(x, "the id tensor")
    if x.dim() != 2:
        raise ValueError(f"expected ids of shape (batch, n_cols), got {tuple(x.shape)}")
    if x.dtype != torch.int32:
        if x.dtype not in (torch.int64, torch.int16, torch.int8, torch.uint8):
            raise TypeError(f"ids must be an integer tensor, got {x.dtype}")
        x = x.to(torch.int32)
    return x.contiguous()
