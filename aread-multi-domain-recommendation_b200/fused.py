"""The whole AREAD forward (lookup -> row pass -> experts -> HEI levels -> heads) as ONE autograd node.

PyTorch sees a single Function whose inputs are the id batch and every parameter on the path and
whose output is the stack of per-tower probabilities; inside, forward and backward are straight
sequences of C-ABI kernel launches (csrc/*.cu) on packed parameter storage.  Nothing here is
traced or compiled; the few torch calls left operate on parameter-sized vectors (assembling the
row-pass weight matrix, the cross-network constants) or allocate buffers.

Reference: model/aread.py:129-322 (forward modes + hier_tower_mask_forward).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from . import _mem
from . import dense_kernels as dk
from . import embedding_ops
from . import hei_ops
from . import rowpass_ops
from . import tower_ops
from .expert_ops import _hi, _lo


# fused tower-layer kernels (csrc/hei.cu); AREAD_HEI_FUSED=0 selects the per-op kernels of tower.cu / bn_act.cu
USE_HEI_LAYER = os.environ.get("AREAD_HEI_FUSED", "1") != "0"


# CUDA-graph replay of the fused node (AREAD_GRAPHS=0 disables).  After GRAPH_AFTER eager calls of one configuration
# (mask, batch shape, train / eval, precision) its forward -- and on the first backward its backward -- launch
# sequence is recorded once and replayed from then on: the activation arena makes every address repeat, the dropout
# seed lives in device memory (dense_kernels.SEED_PTR), the ids are copied into a fixed buffer.
USE_GRAPHS = os.environ.get("AREAD_GRAPHS", "1") != "0"
# skinny products of the row pass (linear / gates / cross / heads) on the tensor cores (AREAD_TC_ROWPASS=0: CUDA cores)
TC_ROWPASS = os.environ.get("AREAD_TC_ROWPASS", "1") != "0"
# BatchNorm-backward masks and column sums inside the data-gradient GEMM epilogue (AREAD_FUSED_BN_BWD=1).  Off by
# default: the epilogue has four warps per SM for elementwise work that wants full occupancy (measured 557 us against
# ~200 us for the GEMM + two streaming passes on the layer-2 gradient at B = 65,536)
FUSED_BN_BWD = os.environ.get("AREAD_FUSED_BN_BWD", "0") == "1"
# device address of the dropout seed while a whole train step is recorded / replayed (step_graph.py)
STEP_SEED_PTR = None
GRAPH_AFTER = 2
# a mask handed in as `current_mask=` is a candidate of the HEMP search: each one is evaluated about
# regroup_eval_step (5) times (run.py:649-655), so recording it costs more than it saves
GRAPH_AFTER_CANDIDATE = 8
MAX_GRAPHS = int(os.environ.get("AREAD_MAX_GRAPHS", "1024"))


class _Stub:
    """Stands in for the autograd context while a launch sequence is recorded."""


class GraphEntry:
    def __init__(self):
        self.calls = 0
        self.fwd = self.bwd = None
        self.x = self.probs = self.ctx = self.d_probs = self.grads = None
        self.fwd_end = 0
        self.sig = None
        self.seed = None                 # int64 [1] device slot: the dropout seed this entry's sequences read
        self.last_use = 0
        self.n_fwd = self.n_bwd = 0      # library launches inside the recorded sequences (for aread_launch_count)


class GraphCache:
    def __init__(self):
        self.entries = {}
        # Two private memory pools: the gradients a recorded backward leaves behind are adopted as .grad and must
        # survive the NEXT forward replay (gradient accumulation), so they may not share blocks with a forward's
        # temporaries.
        self.pool = self.pool_bwd = None
        self.tgrad = {}
        self.clock = 0
        self.force = False              # AREAD.record_graphs: record on the first call instead of after GRAPH_AFTER

    def clear(self):
        self.entries.clear()
        self.tgrad.clear()

    def get(self, key):
        e = self.entries.get(key)
        if e is None:
            if len(self.entries) >= MAX_GRAPHS:          # least recently used goes first
                self.entries.pop(min(self.entries, key=lambda k: self.entries[k].last_use))
            e = self.entries[key] = GraphEntry()
        self.clock += 1
        e.last_use = self.clock
        return e

    def recorded(self, serial, shape, training):
        """True when an entry of the mask with this serial, for batches of this shape in this mode, holds a
        recorded forward."""
        return any(k[0] == serial and k[1] == tuple(shape) and k[2] == training and e.fwd is not None
                   for k, e in self.entries.items())

    def table_grad(self, like):
        """ONE gradient buffer for the embedding table, shared by every recorded backward: a table-sized buffer per
        mask would not fit for large tables, and a recorded gradient only has to live until the optimizer step."""
        key = (like.device, tuple(like.shape))
        t = self.tgrad.get(key)
        if t is None:
            self.tgrad.clear()
            t = self.tgrad[key] = torch.empty_like(like)
        return t



def _signature(model, arena):
    """Everything a recorded sequence has baked in besides its key: the arena buffer and the parameter storage."""
    return (arena.buf_gen, model.embedding.embedding_dict.weight.data_ptr()) + \
        tuple(pk.flat.data_ptr() for pk in model._packs.packs)


class _recording:
    """Stream capture into `graph` without torch.cuda.graph's device synchronisation, garbage collection and
    cache flush (tens of milliseconds each): a side stream that waits for the current one, capture, and back."""

    def __init__(self, graph, pool, device):
        self.graph, self.pool, self.device = graph, pool, device

    def __enter__(self):
        self.side = torch.cuda.Stream(self.device)
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        # thread_local: only this thread is held to the capture rules -- NCCL's watchdog and other helper threads keep
        # polling events while the sequence is recorded
        self.graph.capture_begin(pool=self.pool, capture_error_mode="thread_local")

    def __exit__(self, exc_type, exc, tb):
        try:
            self.graph.capture_end()
        finally:
            self.ctx.__exit__(exc_type, exc, tb)
        if exc_type is None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)


class _seed_ptr:
    def __init__(self, ptr):
        self.ptr = ptr

    def __enter__(self):
        self.prev, dk.SEED_PTR = dk.SEED_PTR, self.ptr

    def __exit__(self, *exc):
        dk.SEED_PTR = self.prev


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _logit_geometry(logits, na, n_prev):
    """(pointer, row stride or 0) of gate logits given as [B, na, n_prev] (contiguous) or as a [B, na * n_prev] column
    slice of a wider matrix."""
    if logits.dim() == 3:
        return logits.data_ptr(), 0
    assert logits.shape[1] == na * n_prev and logits.stride(1) == 1
    return logits.data_ptr(), logits.stride(0)


def gate_mix_fwd(logits, na, n_prev, edges, prev_slot, u_prev, n_prev_active, width, want_sm, sm_escapes=False,
                 logit_offset=None):
    B = logits.shape[0]
    out = _mem.empty((B, na, width), torch.float32, logits.device)
    sm = None
    if want_sm:       # gates handed back to the caller must not live in the arena
        sm = (torch.empty((B, na, n_prev), dtype=torch.float32, device=logits.device) if sm_escapes
              else _mem.empty((B, na, n_prev), torch.float32, logits.device))
    lptr, ld = _logit_geometry(logits, na, n_prev)
    a = _lib.GateMixArgs(B, na, n_prev, n_prev_active, width, lptr,
                         edges.data_ptr() if edges is not None else None, prev_slot.data_ptr(), None,
                         u_prev.data_ptr(), out.data_ptr(), sm.data_ptr() if want_sm else None, None, None, None, None,
                         ld, logit_offset.data_ptr() if logit_offset is not None else None, 0)
    _lib.check(_lib.load().aread_gate_mix(ctypes.byref(a), _stream(logits.device)))
    return out, sm


def gate_mix_bwd(logits, na, n_prev, edges, prev_slot, slot_tower, u_prev, d_out, logit_offset=None, d_logits=None):
    """-> (d_logits, d_u_prev).  `d_logits`: a preallocated [B, na * n_prev] column slice to write into."""
    B = logits.shape[0]
    nap, width = u_prev.shape[1], u_prev.shape[2]
    dev = logits.device
    if d_logits is None:
        d_logits = _mem.empty((B, na, n_prev), torch.float32, dev)
    d_u = _mem.empty((B, nap, width), torch.float32, dev)
    lptr, ld = _logit_geometry(logits, na, n_prev)
    dptr, ldd = _logit_geometry(d_logits, na, n_prev)
    a = _lib.GateMixArgs(B, na, n_prev, nap, width, lptr, edges.data_ptr() if edges is not None else None,
                         prev_slot.data_ptr(), slot_tower.data_ptr(), u_prev.data_ptr(), None, None, d_out.data_ptr(),
                         dptr, d_u.data_ptr(), None, ld,
                         logit_offset.data_ptr() if logit_offset is not None else None, ldd)
    _lib.check(_lib.load().aread_gate_mix(ctypes.byref(a), _stream(dev)))
    return d_logits, d_u


class ModelPacks:
    """Packed views of every parameter family the fused path reads (packing.py)."""

    def __init__(self, model, packs, expert_layers, tower_layers):
        self.experts = expert_layers
        self.towers = tower_layers
        self.mmoe_w = packs.add(lambda: [g[0].weight for g in model.mmoe_gates])          # [n0, NE, E]
        self.mmoe_b = packs.add(lambda: [g[0].bias for g in model.mmoe_gates])            # [n0, NE]
        self.cn_w = packs.add(lambda: [lin.weight for lin in model.cn.w])                 # [nc, 1, E]
        self.cn_b = packs.add(lambda: list(model.cn.b))                                   # [nc, E]
        self.tl_w = packs.add(lambda: [lin.weight for lin in model.towers_linear])        # [n_last, 1, E + w]
        self.gate_w = [None] + [packs.add(lambda l=l: [g[0].weight for g in model.tower_gates[l - 1]])
                                for l in range(1, model.n_level)]                         # [n_l, n_prev, 2D]
        self.gate_b = [None] + [packs.add(lambda l=l: [g[0].bias for g in model.tower_gates[l - 1]])
                                for l in range(1, model.n_level)]


def param_list(model):
    """Every parameter the fused node differentiates, in the order its backward returns gradients."""
    ps = [model.embedding.embedding_dict.weight, model.linear.fc.weight, model.linear.fc.bias,
          model.group_embedding.weight]
    ps += [lin.weight for lin in model.cn.w] + list(model.cn.b)
    ps += [g[0].weight for g in model.mmoe_gates] + [g[0].bias for g in model.mmoe_gates]
    for L in model._expert_layers:
        ps += L.params
    for level in model._tower_layers:
        for L in level:
            ps += L.params
    for l in range(1, model.n_level):
        ps += [g[0].weight for g in model.tower_gates[l - 1]] + [g[0].bias for g in model.tower_gates[l - 1]]
    ps += [lin.weight for lin in model.towers_linear]
    return ps


def _sel(flat, index):
    return flat if index is None else flat.index_select(0, index)


class AreadNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg, *params):
        model = cfg["model"]
        arena = model.arena(x.device)
        ctx.lease = lease = arena.acquire()              # None: another forward of this model awaits its backward
        entry = cfg.get("graph") if lease is not None else None
        ctx.entry = None
        ctx.shards = shards = model.embedding.plan(x.device).shards
        if shards is not None:       # the two collectives of the sharded table stay outside the recorded sequences
            shards.fence()
        if entry is not None:
            sig = _signature(model, arena)
            if entry.fwd is not None and entry.sig != sig:                   # storage moved: record again
                entry.fwd = entry.bwd = None
            record = entry.fwd is None and entry.calls > cfg["graph_after"] and 0 < arena.need <= arena.cap
            if (entry.fwd is not None or record) and cfg["seed"] != 0:
                # the seed travels through device memory so that a recorded sequence sees a new one at every replay.
                # One slot per entry, written only here: the forward that holds the arena lease is the only one whose
                # backward can replay this entry, so nothing overwrites the seed between the two.
                if entry.seed is None:
                    entry.seed = torch.zeros(1, dtype=torch.int64, device=x.device)
                host = torch.empty(1, dtype=torch.int64, pin_memory=True)
                host[0] = cfg["seed"]
                entry.seed.copy_(host, non_blocking=True)
                cfg["seed"], cfg["seed_ptr"] = 0, entry.seed.data_ptr()
            if record:
                AreadNode._record_forward(entry, model, arena, x, cfg, sig)
            if entry.fwd is not None:
                entry.x.copy_(x)
                entry.fwd.replay()
                _lib.load().aread_launch_count_add(entry.n_fwd)
                ctx.entry = entry
                # graph-owned buffers never reach the caller: the next replay would overwrite them
                cfg["gate_means"], cfg["gates"] = {}, {}
                cfg["gate_inputs"] = entry.ctx.cfg["gate_inputs"].clone() if cfg.get("want_gate_inputs") else None
                return entry.probs.clone()
        with _mem.use(arena if lease is not None else None), _seed_ptr(cfg.get("seed_ptr")):
            return AreadNode._forward(ctx, x, cfg)

    @staticmethod
    def _record_forward(entry, model, arena, x, cfg, sig):
        graphs = model._graphs
        if graphs.pool is None:
            graphs.pool, graphs.pool_bwd = torch.cuda.graph_pool_handle(), torch.cuda.graph_pool_handle()
        entry.x = x.clone()
        stub, g = _Stub(), torch.cuda.CUDAGraph()
        arena.off = 0
        n0 = _lib.launch_count()
        with _recording(g, graphs.pool, x.device):
            with _mem.use(arena), _seed_ptr(cfg.get("seed_ptr")):
                probs = AreadNode._forward(stub, entry.x, dict(cfg))
        entry.n_fwd = _lib.launch_count() - n0
        _lib.load().aread_launch_count_add(-entry.n_fwd & 0xFFFFFFFFFFFFFFFF)   # recorded, not run: replay counts them
        entry.fwd, entry.ctx, entry.probs, entry.fwd_end, entry.sig = g, stub, probs, arena.off, sig
        entry.bwd = entry.grads = None

    @staticmethod
    def backward(ctx, d_probs):
        grads = AreadNode._backward_local(ctx, d_probs)
        if ctx.shards is not None:   # owner-major [N * rows, D] buffer -> this rank's averaged shard gradient
            grads = grads[:2] + (ctx.shards.reduce_grad(grads[2]),) + grads[3:]
        return grads

    @staticmethod
    def _backward_local(ctx, d_probs):
        lease = ctx.lease
        if lease is not None and not lease.active:
            raise RuntimeError("the fused AREAD node recycles its activations after the backward; a second backward "
                               "through the same graph (retain_graph=True) needs AREAD_WORKSPACE=0")
        try:
            entry = ctx.entry
            if entry is None:
                with _mem.use(lease.arena if lease is not None else None), _seed_ptr(ctx.cfg.get("seed_ptr")):
                    return AreadNode._backward(ctx, d_probs)
            arena, inner = lease.arena, entry.ctx
            arena.off = entry.fwd_end                    # the backward's temporaries follow the forward's
            # the recorded backward writes its gradients into fixed buffers which autograd then adopts as .grad:
            # only sound when no earlier gradient is being accumulated into
            if all(p.grad is None for p in inner.cfg["model"]._fused_params):
                if entry.bwd is None:
                    entry.d_probs = d_probs.contiguous().clone()
                    g = torch.cuda.CUDAGraph()
                    n0 = _lib.launch_count()
                    model = inner.cfg["model"]
                    if ctx.shards is None:       # (sharded: the owner-major buffer is already one per model)
                        inner.cfg["table_grad"] = model._graphs.table_grad(model.embedding.embedding_dict.weight)
                    try:
                        with _recording(g, model._graphs.pool_bwd, d_probs.device):
                            with _mem.use(arena), _seed_ptr(inner.cfg.get("seed_ptr")):
                                out = AreadNode._backward(inner, entry.d_probs)
                    finally:
                        inner.cfg.pop("table_grad", None)    # eager launches keep allocating their own
                    entry.bwd, entry.grads, entry.n_bwd = g, out[2:], _lib.launch_count() - n0
                    _lib.load().aread_launch_count_add(-entry.n_bwd & 0xFFFFFFFFFFFFFFFF)
                entry.d_probs.copy_(d_probs)
                entry.bwd.replay()
                _lib.load().aread_launch_count_add(entry.n_bwd)
                return (None, None, *[None if t is None else t.detach() for t in entry.grads])
            with _mem.use(arena), _seed_ptr(inner.cfg.get("seed_ptr")):
                return AreadNode._backward(inner, d_probs)
        finally:
            if lease is not None:
                lease.release()

    @staticmethod
    def _forward(ctx, x, cfg):
        model, info = cfg["model"], cfg["info"]
        P = model._fused
        dev = x.device
        training, p_drop, precise = model.training, model.dropout_p, cfg["precise"]
        seed = cfg["seed"]
        n_level, n_tower = model.n_level, model.n_tower
        D, E = model.embed_dim, model.embed_output_dim
        B = x.shape[0]
        bn_skip = B == 1
        sv = {}                                                     # tensors kept for the backward

        # ---- lookup
        plan = model.embedding.plan(dev)
        table = model.embedding.embedding_dict.weight
        embed, xb = embedding_ops.gather(plan, table, x, want_bf16=True, want_lo=True, fence=False)
        X = embed.view(B, E)
        x_hi, x_lo = xb                     # X = hi + lo (bf16 each): operands of the tensor-core products
        if not precise:
            xb = x_hi

        # ---- which towers run
        active = [list(range(n)) for n in n_tower] if info is None else info.active_idx
        index = [None if (info is None or len(active[l]) == n_tower[l]) else info.index(l, dev)
                 for l in range(n_level)]
        a0, a_last = active[0], active[-1]
        n_expert = len(model.mmoe_experts)
        n_cross = model.cn.num_layers

        # ---- mean group embedding of the active level-0 towers (second half of the HEI gate input q)
        if info is None:
            grp = torch.zeros(D, dtype=torch.float32, device=dev)
        else:
            grp = model.group_embedding.weight.index_select(0, info.group_index(dev)).mean(dim=0)

        # ---- row pass: [linear | gates of active level-0 towers | cross | heads of active last-level towers] and,
        # riding behind them, the HEI gate logits: q = [X[:, domain field] | grp], so W_gate . q is a dot product of
        # the embedding row (weights outside the domain field's columns are zero) plus a row-independent constant
        w_out = _sel(P.tl_w.flat, index[-1])[:, 0, :]                                  # [na_last, E + w]
        w_rows = [model.linear.fc.weight, _sel(P.mmoe_w.flat, index[0]).reshape(-1, E),
                  P.cn_w.flat.view(n_cross, E), w_out[:, :E]]
        beta = torch.cumsum(P.cn_b.flat, dim=0)                                          # beta_{k+1} = b_0 + .. + b_k
        kappa = torch.zeros(n_cross, dtype=torch.float32, device=dev)
        if n_cross > 1:
            kappa[1:] = (P.cn_w.flat.view(n_cross, E)[1:] * beta[:-1]).sum(dim=1)
        beta_n = beta[-1] if n_cross > 0 else torch.zeros(E, dtype=torch.float32, device=dev)
        offs = [model.linear.fc.bias, _sel(P.mmoe_b.flat, index[0]).reshape(-1), kappa, w_out[:, :E] @ beta_n]
        layout = (len(a0), n_expert, n_cross, len(a_last))
        nj_own = 1 + len(a0) * n_expert + n_cross + len(a_last)
        n_gate_logits = sum(len(active[l]) * n_tower[l - 1] for l in range(1, n_level))
        # ('bf16x3' is the tight-parity mode: it keeps the all-fp32 CUDA-core row pass -- the split operands carry
        # 16-17 bits, not 24, and this model amplifies forward perturbations into the gradients behind the towers)
        tc_row = TC_ROWPASS and not precise and E % 8 == 0 and nj_own <= 128
        ride = tc_row and n_gate_logits > 0 and nj_own + n_gate_logits <= 128
        gate_cols = [None] * n_level                                                     # (first column, wg) per level
        if ride:
            dom0, col = model.domain_idx * D, nj_own
            for l in range(1, n_level):
                wg = _sel(P.gate_w[l].flat, index[l])                                     # [na, n_prev, 2D]
                bg = _sel(P.gate_b[l].flat, index[l])
                n_rows = wg.shape[0] * wg.shape[1]
                w_rows.append(torch.nn.functional.pad(wg[:, :, :D].reshape(n_rows, D), (dom0, E - dom0 - D)))
                offs.append(bg.reshape(-1) + wg[:, :, D:].reshape(n_rows, D) @ grp)
                gate_cols[l] = (col, wg)
                col += n_rows
        w_cat = torch.cat(w_rows, dim=0)
        offset = torch.cat(offs, dim=0)
        nj = w_cat.shape[0]
        ldp = (nj + 3) // 4 * 4
        p_dots = _mem.empty((B, ldp), torch.float32, dev)
        lin = _mem.empty((B,), torch.float32, dev)
        gate = _mem.empty((B, len(a0), n_expert), torch.float32, dev)
        alpha = _mem.empty((B, n_cross + 1), torch.float32, dev)
        head_cross = _mem.empty((B, len(a_last)), torch.float32, dev)
        # the [B, E] x [E, nj] product on the tensor cores with split operands (hi.hi + hi.lo + lo.hi + lo.lo into one
        # fp32 accumulator: the full product of the split values), the per-row epilogue on the CUDA cores
        # (AREAD_TC_ROWPASS=0 keeps the all-fp32 CUDA-core kernels of rowpass.cu)
        if tc_row:
            wc_hi, wc_lo = dk.split_bf16(w_cat)
            dk.grouped_linear(x_hi, wc_hi, None, nj, E, 1, 0, out=p_dots, a_lo=x_lo, w_lo=wc_lo, lo_lo=True)
            sv.update(wc_hi=wc_hi, wc_lo=wc_lo, x_hi=x_hi, x_lo=x_lo)
        ra = rowpass_ops._args(B, E, layout, ldp, x=None if tc_row else X, w=None if tc_row else w_cat, offset=offset,
                               p=p_dots, lin=lin, gate=gate, alpha=alpha, head=head_cross)
        ra.n_extra = nj - nj_own
        _lib.check(_lib.load().aread_rowpass_fwd(ctypes.byref(ra), _stream(dev)))
        sv.update(X=X, w_cat=w_cat, p_dots=p_dots, gate=gate, alpha=alpha, beta=beta, w_out=w_out, layout=layout, ldp=ldp,
                  tc_row=tc_row, ride=ride, nj_own=nj_own, offset=offset, gate_cols=gate_cols, grp=grp)

        # ---- experts (tensor cores) + MMoE mixture
        G = P.experts[0].groups
        a_op = xb
        ex = []
        counters = []       # num_batches_tracked of every BatchNorm that ran in train mode: ONE increment launch at the end
        # the fused epilogues store 64-column bf16 chunks: narrower expert layers keep the unfused kernels
        fused_bn = not precise and all(L.n % 64 == 0 for L in P.experts)
        sv["fused_bn"] = fused_bn
        if not fused_bn:
            for i, L in enumerate(P.experts):
                w = dk.split_bf16(L.weight.rows()) if precise else L.weight.rows().to(torch.bfloat16)
                z = dk.grouped_linear(_hi(a_op), _hi(w), L.bias.rows(), L.n, L.k, G, 0 if i == 0 else L.k,
                                      a_lo=_lo(a_op), w_lo=_lo(w))
                last = i == len(P.experts) - 1
                res = dk.bn_act_fwd(z, L.gamma.rows(), L.beta.rows(), L.running_mean.rows(), L.running_var.rows(),
                                    training, bn_skip, p_drop, seed, L.salt, None if last else torch.bfloat16,
                                    want_lo=precise and not last)
                out, stats = (res[0], res[-1]) if not (precise and not last) else ((res[0], res[1]), res[2])
                if training and not bn_skip:
                    counters.extend(L.tracked)
                ex.append((a_op, z, stats, w, None))
                a_op = out
        else:
            # bf16 experts: the GEMM epilogue leaves the BatchNorm column sums and the pre-activation ONCE, as bf16
            # (bias-free: BatchNorm removes it); under no_grad in eval mode BatchNorm + ReLU are folded into the epilogue
            w16 = model.expert_weights_bf16()
            keep_z = training or torch.is_grad_enabled()
            for i, L in enumerate(P.experts):
                last = i == len(P.experts) - 1
                agc = 0 if i == 0 else L.k
                bn = (L.bias.rows(), L.gamma.rows(), L.beta.rows(), L.running_mean.rows(), L.running_var.rows())
                if keep_z:
                    z, partial = dk.expert_linear_stats(a_op, w16[i], L.n, L.k, G, agc)
                    stats = dk.expert_bn_finalize(partial, B, G * L.n, *bn, training, bn_skip)
                    out = bits = None
                    if not last:      # with a backward to come, the ReLU / dropout pattern is kept as one bit per element
                        want = torch.is_grad_enabled() and not FUSED_BN_BWD
                        res = dk.bn16_fwd(z, stats, training, p_drop, seed, L.salt, want_bits=want)
                        out, bits = res if want else (res, None)
                else:
                    folded = dk.expert_bn_finalize(None, B, G * L.n, *bn, False, bn_skip)
                    out = z = dk.expert_linear_act(a_op, w16[i], L.n, L.k, G, agc, folded)
                    stats = dk.identity_saved(G * L.n, dev) if last else None
                    bits = None
                if training and not bn_skip:
                    counters.extend(L.tracked)
                ex.append((a_op, z, stats, w16[i], bits if keep_z else None))
                a_op = out
        h = dk.mmoe_mix_fwd(ex[-1][1], ex[-1][2], gate, G, len(a0), p_drop if training else 0.0, seed,
                            P.experts[-1].salt)                                          # [B, na0, H]
        sv["experts"] = ex

        # ---- gate inputs q = [domain embedding | mean group embedding of the active level-0 towers]
        q = None
        if not ride:
            q = torch.cat([embed[:, model.domain_idx, :], grp.expand(B, D)], dim=1)      # [B, 2D]
        sv["q"] = q

        # ---- HEI levels on compact activations
        gate_means, gates = {}, {}
        levels = []
        for l in range(n_level):
            act, idx = active[l], index[l]
            rec = {}
            if l > 0:
                n_prev = n_tower[l - 1]
                logit_offset = None
                if ride:        # the logits are columns of the row-pass product; the constants are added on the fly
                    col, wg = gate_cols[l]
                    logits = p_dots[:, col:col + len(act) * n_prev]
                    logit_offset = offset[col:col + len(act) * n_prev]
                else:
                    wg, bg = _sel(P.gate_w[l].flat, idx), _sel(P.gate_b[l].flat, idx)
                    # all gates of the level share the input q: ONE [B, 2D] x [2D, na * n_prev] product
                    if len(act) * n_prev <= 128:
                        logits = tower_ops.tower_linear(q, wg.view(1, len(act) * n_prev, -1), bg.reshape(-1),
                                                        len(act) * n_prev, groups=1).view(B, len(act), n_prev)
                    else:
                        logits = tower_ops.tower_linear(q, wg, bg.reshape(-1), n_prev, groups=len(act))
                edges = None if info is None else _sel(info.edges(l, dev).t().contiguous(), idx)
                prev_slot, slot_tower = cfg["slots"][l]
                want_sm = cfg["want_gate_means"] or cfg["want_gates"]
                u_prev = h
                h, sm = gate_mix_fwd(logits, len(act), n_prev, edges, prev_slot, u_prev, len(active[l - 1]),
                                     u_prev.shape[2], want_sm, sm_escapes=cfg["want_gates"], logit_offset=logit_offset)
                if cfg["want_gates"]:
                    gates[l] = sm.transpose(1, 2)                                        # [B, n_prev, n_l]
                if cfg["want_gate_means"] and info is not None:
                    means = sm.mean(dim=0).t()                                           # [n_prev, na]
                    if idx is not None:     # towers that do not run report zeros (aread.py:278-280)
                        means = torch.zeros(n_prev, n_tower[l], dtype=torch.float32,
                                            device=dev).index_copy_(1, idx, means)
                    gate_means[l] = means
                rec.update(wg=wg, logits=logits, logit_offset=logit_offset, edges=edges, u_prev=u_prev)
            lay = []
            na = len(act)
            hei = USE_HEI_LAYER and all(hei_ops.supported(na, L.k, L.n) for L in P.towers[l])
            rec["hei"] = hei
            z = stats = None
            for j, L in enumerate(P.towers[l]):
                full = idx is None
                w, bias = _sel(L.weight.flat, idx), _sel(L.bias.flat, idx)
                gamma, beta_bn = _sel(L.gamma.flat, idx), _sel(L.beta.flat, idx)
                rm, rv = _sel(L.running_mean.flat, idx), _sel(L.running_var.flat, idx)
                if hei:     # the layer reads its input straight from the previous pre-activation
                    src = h.view(B, na * L.k) if j == 0 else z
                    src_saved = None if j == 0 else stats
                    src_salt = 0 if j == 0 else P.towers[l][j - 1].salt
                    z, stats = hei_ops.layer_fwd(src, src_saved, src_salt, w, bias, gamma.reshape(-1), beta_bn.reshape(-1),
                                                 rm.view(-1), rv.view(-1), na, L.k, L.n, training, bn_skip, p_drop, seed)
                    lay.append((src, src_saved, z, stats, w))
                else:
                    z = tower_ops.tower_linear(h, w, bias, L.n)
                    out, stats = dk.bn_act_fwd(z.view(B, na * L.n), gamma.reshape(-1), beta_bn.reshape(-1), rm.view(-1),
                                               rv.view(-1), training, bn_skip, p_drop, seed, L.salt, torch.float32)
                    lay.append((h, z, stats, w))
                    h = out.view(B, na, L.n)
                if training and not bn_skip:
                    if not full:
                        L.running_mean.flat.index_copy_(0, idx, rm)
                        L.running_var.flat.index_copy_(0, idx, rv)
                    tracked = L.tracked
                    counters.extend(tracked[t] for t in act)
            if hei:
                L = P.towers[l][-1]
                h = hei_ops.bn_apply(z, stats, training, p_drop, seed, L.salt).view(B, na, L.n)
            rec["layers"] = lay
            levels.append(rec)
        sv["levels"] = levels
        if counters:
            torch._foreach_add_(counters, 1)

        # ---- heads: z_t = w_out_t[:E] . cn_out + w_out_t[E:] . u_t + lin ; p = sigmoid(z)
        w_tail = w_out[:, E:].contiguous()                                               # [na_last, w]
        probs = torch.empty((len(a_last), B), dtype=torch.float32, device=dev)           # leaves the node: not arena
        ha = _lib.HeadArgs(B, len(a_last), w_tail.shape[1], head_cross.data_ptr(), lin.data_ptr(), h.data_ptr(),
                           w_tail.data_ptr(), probs.data_ptr(), None, None, None, None, None, None, 0)
        _lib.check(_lib.load().aread_head(ctypes.byref(ha), _stream(dev)))
        sv.update(h_last=h, w_tail=w_tail, probs=probs.detach())      # detached alias: no ctx <-> output cycle

        if q is None and cfg.get("want_gate_inputs"):
            q = torch.cat([embed[:, model.domain_idx, :], grp.expand(B, D)], dim=1)
        cfg["gate_means"], cfg["gates"], cfg["gate_inputs"] = gate_means, gates, q
        ctx.cfg, ctx.sv = cfg, sv
        ctx.active, ctx.index, ctx.bn_skip = active, index, bn_skip
        ctx.x_ids = x
        return probs

    @staticmethod
    def _backward(ctx, d_probs):
        cfg, sv = ctx.cfg, ctx.sv
        model, info = cfg["model"], cfg["info"]
        P = model._fused
        active, index, bn_skip = ctx.active, ctx.index, ctx.bn_skip
        training, precise, seed = model.training, cfg["precise"], cfg["seed"]
        p_drop = model.dropout_p if training else 0.0
        n_level, n_tower = model.n_level, model.n_tower
        D, E = model.embed_dim, model.embed_output_dim
        X = sv["X"]
        B = X.shape[0]
        dev = X.device
        n_cross = model.cn.num_layers
        n_expert = len(model.mmoe_experts)

        probs = sv["probs"]
        na_last, w_last = probs.shape[0], sv["w_tail"].shape[1]
        dz = _mem.empty((B, na_last), torch.float32, dev)
        d_lin = _mem.empty((B,), torch.float32, dev)
        d_h = _mem.empty((B, na_last, w_last), torch.float32, dev)
        d_w_tail = torch.empty((na_last, w_last), dtype=torch.float32, device=dev)       # parameter gradient
        need = int(_lib.load().aread_head_workspace_bytes(na_last, w_last))
        ws = _mem.workspace("head", dev, need)
        ha = _lib.HeadArgs(B, na_last, w_last, None, None, sv["h_last"].data_ptr(), sv["w_tail"].data_ptr(),
                           probs.data_ptr(), d_probs.contiguous().data_ptr(), dz.data_ptr(), d_lin.data_ptr(),
                           d_h.data_ptr(), d_w_tail.data_ptr(), ws.data_ptr(), ws.numel())
        _lib.check(_lib.load().aread_head(ctypes.byref(ha), _stream(dev)))

        layout, ldp, tc_row, ride = sv["layout"], sv["ldp"], sv["tc_row"], sv["ride"]
        nj, nj_own = sv["w_cat"].shape[0], sv["nj_own"]
        d_p = _mem.empty((B, ldp), torch.float32, dev)     # gradient w.r.t. the row-pass products (gate logits included)
        # only the column sums of d_c (gradients of the additive constants) are needed: the prologue kernel adds them up
        d_off_full = torch.empty((ldp,), dtype=torch.float32, device=dev)       # parameter gradients: never arena
        dc_partial = _mem.workspace("rowpass_dc", dev, int(_lib.load().aread_rowpass_prologue_ctas(B)) * ldp * 4)

        tower_grads = [[None] * len(P.towers[l]) for l in range(n_level)]
        gate_grads = [None] * n_level
        d_q = None
        for l in range(n_level - 1, -1, -1):
            rec = sv["levels"][l]
            na = len(active[l])
            if rec["hei"]:
                layers = P.towers[l]
                L = layers[-1]
                d_out = d_h.contiguous().view(B, na * L.n)
                coef, g3 = hei_ops.bn_bwd_coef(rec["layers"][-1][2], d_out, rec["layers"][-1][3], bn_skip, p_drop, seed,
                                               L.salt)
                for j in range(len(layers) - 1, -1, -1):
                    L = layers[j]
                    src, src_saved, z, stats, w = rec["layers"][j]
                    d_out, d_w, src_coef, src_g3 = hei_ops.layer_bwd(z, d_out, stats, coef, p_drop, L.salt, seed, bn_skip,
                                                                     src, src_saved, layers[j - 1].salt if j > 0 else 0,
                                                                     w, na, L.k, L.n)
                    tower_grads[l][j] = (d_w, g3[2].view(na, L.n), g3[0].view(na, L.n), g3[1].view(na, L.n))
                    coef, g3 = src_coef, src_g3
                d_h = d_out.view(B, na, layers[0].k)
            for j in range(len(P.towers[l]) - 1, -1, -1) if not rec["hei"] else ():
                L = P.towers[l][j]
                h_in, z, stats, w = rec["layers"][j]
                dzl, d_gamma, d_beta, d_bias = dk.bn_act_bwd(z.view(B, na * L.n), d_h.contiguous().view(B, na * L.n),
                                                             stats, bn_skip, p_drop, seed, L.salt, torch.float32)
                dzl = dzl.view(B, na, L.n)
                d_w = tower_ops.tower_wgrad(dzl, h_in)
                tower_grads[l][j] = (d_w, d_bias.view(na, L.n), d_gamma.view(na, L.n), d_beta.view(na, L.n))
                d_h = tower_ops.tower_linear(dzl, w, None, L.k, weight_is_out_by_in=False)
            if l > 0:
                prev_slot, slot_tower = cfg["slots"][l]
                n_prev = n_tower[l - 1]
                if ride:    # d_logits are columns of d_p: weight / input gradients come out of the row pass products
                    col = sv["gate_cols"][l][0]
                    _, d_u = gate_mix_bwd(rec["logits"], na, n_prev, rec["edges"], prev_slot, slot_tower, rec["u_prev"],
                                          d_h.contiguous(), logit_offset=rec["logit_offset"],
                                          d_logits=d_p[:, col:col + na * n_prev])
                else:
                    d_logits, d_u = gate_mix_bwd(rec["logits"], na, n_prev, rec["edges"], prev_slot, slot_tower,
                                                 rec["u_prev"], d_h.contiguous())
                    if ((na * n_prev + 3) // 4) * ((2 * D + 3) // 4) <= 256:             # one group: q is read once
                        d_wg = tower_ops.tower_wgrad(d_logits.view(B, 1, na * n_prev), sv["q"]).view(na, n_prev, 2 * D)
                    else:
                        d_wg = tower_ops.tower_wgrad(d_logits, sv["q"])                  # [na, n_prev, 2D]
                    d_bg = d_logits.sum(dim=0)                                           # [na, n_prev]
                    gate_grads[l] = (d_wg, d_bg)
                    dq_l = tower_ops.tower_linear(d_logits.view(B, 1, na * n_prev),
                                                  rec["wg"].reshape(1, na * n_prev, 2 * D), None, 2 * D,
                                                  weight_is_out_by_in=False).view(B, 2 * D)
                    d_q = dq_l if d_q is None else d_q + dq_l
                d_h = d_u

        # ---- experts
        ex = sv["experts"]
        G = P.experts[0].groups
        na0 = len(active[0])
        d_act, d_gate = dk.mmoe_mix_bwd(ex[-1][1], ex[-1][2], sv["gate"], d_h.contiguous(), G, na0, p_drop, seed,
                                        P.experts[-1].salt)
        expert_grads = [None] * len(P.experts)
        d_x = None
        L0 = P.experts[0]
        fused_bn = sv["fused_bn"]
        w16 = (nj + 31) // 32 * 32       # columns of each third of the split row-pass gradient
        c0 = G * L0.n if fused_bn else 0 # with the fused expert path the split rides behind the layer-1 gradient
        k_ext = c0 + 3 * w16             # [B, 4*256 | hi | hi | lo]
        dz0 = None
        if tc_row:
            # per-row prologue of the row pass now.  Its d_p becomes split bf16 operands [hi | hi | lo] which, against
            # the weight rows [hi ; lo ; hi], give d_x of the row pass on the tensor cores -- as extra reduction columns
            # of the expert layer-1 data gradient GEMM when that runs in the fused bf16 path
            dz0 = _mem.empty((B, k_ext), torch.bfloat16, dev)
            ra = rowpass_ops._args(B, E, layout, ldp, x=None, p=sv["p_dots"], gate=sv["gate"], alpha=sv["alpha"],
                                   d_lin=d_lin, d_gate=d_gate, d_head=dz, d_p=d_p, d_c_sum=d_off_full,
                                   d_c_partial=dc_partial)
            ra.dp16, ra.ld16, ra.n_extra, ra.dp16_width = dz0.data_ptr() + c0 * 2, k_ext, nj - nj_own, w16
            _lib.check(_lib.load().aread_rowpass_bwd(ctypes.byref(ra), _stream(dev)))

        def split_weight_rows(w_ext, base):
            """rows [Wcat_hi ; Wcat_lo ; Wcat_hi] (w16 each, zero padded) of the data-gradient weight operand"""
            w_ext[base:base + nj].copy_(sv["wc_hi"])
            w_ext[base + w16:base + w16 + nj].copy_(sv["wc_lo"])
            w_ext[base + 2 * w16:base + 2 * w16 + nj].copy_(sv["wc_hi"])
            if nj < w16:
                for o in (0, w16, 2 * w16):
                    w_ext[base + o + nj:base + o + w16].zero_()

        def tr(w, fn):
            return tuple(fn(t) for t in w) if isinstance(w, tuple) else fn(w)

        if not sv["fused_bn"]:
            for i in range(len(P.experts) - 1, -1, -1):
                L = P.experts[i]
                a_in, z, stats, w, _ = ex[i]
                dze, d_gamma, d_beta, d_bias = dk.bn_act_bwd(z, d_act, stats, bn_skip, p_drop, seed, L.salt,
                                                             want_lo=precise)
                d_w = dk.grouped_wgrad(_hi(dze), _hi(a_in), L.n, L.k, G, 0 if i == 0 else L.k, dz_lo=_lo(dze),
                                       a_lo=_lo(a_in))
                expert_grads[i] = (d_w.view(G, L.n, L.k), d_bias.view(G, L.n), d_gamma.view(G, L.n), d_beta.view(G, L.n))
                if i > 0:
                    wt = tr(w, lambda t: t.view(G, L.n, L.k).transpose(1, 2).reshape(G * L.k, L.n).contiguous())
                    d_act = dk.grouped_linear(_hi(dze), _hi(wt), None, L.k, L.n, G, L.n, a_lo=_lo(dze), w_lo=_lo(wt))
                else:
                    wt = tr(w, lambda t: t.t().contiguous())
                    d_x = dk.grouped_linear(_hi(dze), _hi(wt), None, L.k, G * L.n, 1, 0, a_lo=_lo(dze), w_lo=_lo(wt))
        else:
            # bf16 experts.  Last layer: its gradient arrives in fp32 from the mixture.  Every other layer: the data
            # gradient GEMM of the layer above masks it (ReLU, dropout), sums the BatchNorm reductions in its epilogue
            # and stores it once as bf16; the weight is read in place as a k-by-n operand (no transposed copies).
            n_l = len(P.experts)
            L = P.experts[-1]
            out0 = dz0[:, :G * L0.n] if (tc_row and n_l == 1) else None
            dze, d_gamma, d_beta, d_bias = dk.bn_act_bwd(ex[-1][1], d_act, ex[-1][2], bn_skip, p_drop, seed, L.salt,
                                                         dz_out=out0)
            for i in range(n_l - 1, -1, -1):
                L = P.experts[i]
                a_in, z, stats, w, _ = ex[i]
                d_w = dk.grouped_wgrad(dze, a_in, L.n, L.k, G, 0 if i == 0 else L.k)
                expert_grads[i] = (d_w.view(G, L.n, L.k), d_bias.view(G, L.n), d_gamma.view(G, L.n), d_beta.view(G, L.n))
                if i > 0:
                    Lp = P.experts[i - 1]
                    _, z_p, stats_p, _, bits_p = ex[i - 1]
                    out_l = dz0[:, :G * L0.n] if (tc_row and i == 1) else None
                    if FUSED_BN_BWD:    # masks + BatchNorm sums in the GEMM epilogue (4 epilogue warps do the elementwise work)
                        dy, partial = dk.expert_dgrad_bn_bwd(dze, w, L.k, L.n, G, z_p, stats_p, p_drop, Lp.salt, seed)
                        coef, g3 = dk.expert_bn_bwd_finalize(partial, B, G * L.k, bn_skip)
                        dze = dk.bn16_bwd(z_p, dy, stats_p, coef, bn_skip, out=out_l)
                    else:               # plain bf16 data gradient, then two full-occupancy passes over (z16, d_h16)
                        d_h16 = dk.expert_dgrad_bf16(dze, w, L.k, L.n, G)
                        partial = dk.bn16_bwd_stats(z_p, d_h16, stats_p, bn_skip, p_drop, Lp.salt, seed, bits=bits_p)
                        coef, g3 = dk.expert_bn_bwd_finalize(partial, B, G * L.k, bn_skip)
                        dze = dk.bn16_bwd(z_p, d_h16, stats_p, coef, bn_skip, out=out_l, raw=True, p=p_drop, salt=Lp.salt,
                                          seed=seed, bits=bits_p)
                    d_gamma, d_beta, d_bias = g3[0], g3[1], g3[2]
                elif tc_row:
                    # weight rows [W_layer1 (k-by-n) ; Wcat_hi ; Wcat_lo ; Wcat_hi]: ONE GEMM returns d_x of both paths
                    w_ext = _mem.empty((k_ext, E), torch.bfloat16, dev)
                    w_ext[:G * L.n].copy_(w)
                    split_weight_rows(w_ext, G * L.n)
                    d_x = dk.expert_dgrad_plain(dz0, w_ext, E, k_ext)
                else:
                    d_x = dk.expert_dgrad_plain(dze, w, L.k, G * L.n)

        # ---- row pass
        if tc_row:
            # d_w[j, :] = sum_b d_p[b, j] X[b, :] on the tensor cores, split operands on both sides
            dp_hi, dp_lo = dz0[:, c0:c0 + w16], dz0[:, c0 + 2 * w16:c0 + 3 * w16]
            d_wcat = dk.grouped_wgrad(dp_hi, sv["x_hi"], w16, E, 1, 0, dz_lo=dp_lo, a_lo=sv["x_lo"])[:nj]
            if not fused_bn:     # the expert data gradient ran on its own: the row pass adds its part
                w_ext = _mem.empty((k_ext, E), torch.bfloat16, dev)
                split_weight_rows(w_ext, 0)
                d_x = d_x.add_(dk.expert_dgrad_plain(dz0, w_ext, E, k_ext))
        else:
            d_x_row = _mem.empty((B, E), torch.float32, dev)
            d_wcat = torch.empty((nj, E), dtype=torch.float32, device=dev)                   # parameter gradients
            need = int(_lib.load().aread_rowpass_workspace_bytes(B, E, nj))
            ws = _mem.workspace("rowpass", dev, need)
            ra = rowpass_ops._args(B, E, layout, ldp, x=X, w=sv["w_cat"], p=sv["p_dots"], gate=sv["gate"],
                                   alpha=sv["alpha"], d_lin=d_lin, d_gate=d_gate, d_head=dz, d_p=d_p,
                                   d_c_sum=d_off_full, d_c_partial=dc_partial, d_x=d_x_row, d_w=d_wcat, workspace=ws)
            ra.workspace_bytes = ws.numel()
            _lib.check(_lib.load().aread_rowpass_bwd(ctypes.byref(ra), _stream(dev)))
            d_x = d_x.add_(d_x_row)
        d_off = d_off_full[:nj]
        d_grp_vec = None if d_q is None else d_q[:, D:].sum(dim=0)          # gradient w.r.t. the mean group embedding
        if ride:
            # gate parameters from the riding columns: W[:, :D] from the weight gradient rows (domain field columns),
            # W[:, D:] and the bias from the column sums (the group half of q is the same for every row)
            dom0, grp = model.domain_idx * D, sv["grp"]
            d_grp_vec = torch.zeros(D, dtype=torch.float32, device=dev)
            for l in range(1, n_level):
                col, wg = sv["gate_cols"][l]
                na, n_prev = wg.shape[0], wg.shape[1]
                seg = d_off[col:col + na * n_prev]
                d_wg = torch.cat([d_wcat[col:col + na * n_prev, dom0:dom0 + D], seg[:, None] * grp[None, :]], dim=1)
                gate_grads[l] = (d_wg.view(na, n_prev, 2 * D), seg.view(na, n_prev))
                d_grp_vec = d_grp_vec + seg @ wg[:, :, D:].reshape(na * n_prev, D)

        # ---- gradient w.r.t. the embedding output -> table
        if d_q is not None:
            d_x.view(B, -1, D)[:, model.domain_idx, :] += d_q[:, :D]
        plan = model.embedding.plan(dev)
        d_table = embedding_ops.scatter(plan, ctx.x_ids, d_x, d_table=cfg.get("table_grad"),
                                        reduce=False)            # sharded: owner-major, reduced by the caller

        # ---- unpack d_wcat / d_off into parameter gradients
        ng = na0 * n_expert
        c_cross, c_head = 1 + ng, 1 + ng + n_cross
        d_lin_w, d_lin_b = d_wcat[0:1], d_off[0:1]
        d_mmoe_w, d_mmoe_b = d_wcat[1:c_cross].view(na0, n_expert, E), d_off[1:c_cross].view(na0, n_expert)
        d_cn_w = d_wcat[c_cross:c_head].clone()                                          # [nc, E]
        d_head_w = d_wcat[c_head:nj_own]                                                 # [na_last, E]
        d_kappa, d_rho = d_off[c_cross:c_head], d_off[c_head:nj_own]
        beta, w_out = sv["beta"], sv["w_out"]
        cn_w = P.cn_w.flat.view(n_cross, E)
        # kappa_k = w_k . beta_{k-1}(cum) ; rho_t = w_out_t[:E] . beta_n
        d_beta = torch.zeros(n_cross, E, dtype=torch.float32, device=dev)               # d/d beta_cum[k]
        if n_cross > 1:
            d_cn_w[1:] += d_kappa[1:, None] * beta[:-1]
            d_beta[:-1] += d_kappa[1:, None] * cn_w[1:]
        if n_cross > 0:
            d_beta[-1] += d_rho @ w_out[:, :E]
            d_head_w = d_head_w + d_rho[:, None] * beta[-1]
        d_cn_b = torch.flip(torch.cumsum(torch.flip(d_beta, dims=[0]), dim=0), dims=[0])  # beta_cum[k] = sum_{i<=k} b_i
        d_tl = torch.cat([d_head_w, d_w_tail], dim=1)                                    # [na_last, E + w]

        # ---- group embedding
        d_grp_w = None
        if info is not None and d_grp_vec is not None:
            d_grp_w = torch.zeros_like(model.group_embedding.weight)
            gi = info.group_index(dev)
            d_grp_w.index_add_(0, gi, (d_grp_vec / gi.numel()).expand(gi.numel(), D))
        elif info is not None:
            d_grp_w = torch.zeros_like(model.group_embedding.weight)

        # ---- assemble in param_list order
        def scatter_rows(n_total, act, rows):
            out = [None] * n_total
            for t, row in zip(act, rows.unbind(0)):
                out[t] = row
            return out

        grads = [d_table, d_lin_w, d_lin_b, d_grp_w]
        grads += [d_cn_w[k:k + 1] for k in range(n_cross)] + [d_cn_b[k] for k in range(n_cross)]
        grads += scatter_rows(n_tower[0], active[0], d_mmoe_w) + scatter_rows(n_tower[0], active[0], d_mmoe_b)
        for g4 in expert_grads:
            for tens in g4:
                grads += list(tens.unbind(0))
        for l in range(n_level):
            for j in range(len(P.towers[l])):
                for tens in tower_grads[l][j]:
                    grads += scatter_rows(n_tower[l], active[l], tens)
        for l in range(1, n_level):
            grads += scatter_rows(n_tower[l], active[l], gate_grads[l][0]) + \
                scatter_rows(n_tower[l], active[l], gate_grads[l][1])
        grads += scatter_rows(n_tower[-1], active[-1], d_tl.unsqueeze(1))
        return (None, None, *grads)


def slot_maps(active_prev, n_prev, device, cache):
    """(prev_slot [n_prev] int32: compact index of previous-level tower j or -1, slot_tower [n_active])."""
    key = (tuple(active_prev), n_prev, device)
    got = cache.get(key)
    if got is None:
        slot = np.full(n_prev, -1, dtype=np.int32)
        slot[np.asarray(active_prev, dtype=np.int64)] = np.arange(len(active_prev), dtype=np.int32)
        got = (torch.from_numpy(slot).to(device), torch.tensor(active_prev, dtype=torch.int32, device=device))
        cache[key] = got
    return got


def forward(model, x, info, want_gate_means=False, want_gates=False, want_gate_inputs=False, may_record=True):
    """Runs the fused node; returns (probs [n_active_last, B], cfg) where cfg carries the side outputs.
    `may_record=False` (candidate masks of the HEMP search, which are evaluated a handful of times) raises the
    number of eager calls before a CUDA graph is recorded from GRAPH_AFTER to GRAPH_AFTER_CANDIDATE."""
    table = model.embedding.embedding_dict.weight
    if model._fused_params[0] is not table:        # the table parameter was replaced (shard_table, load)
        model._fused_params[0] = table
    x = embedding_ops.prepare_ids(x, table)
    dev = x.device
    training = model.training
    n_level, n_tower = model.n_level, model.n_tower
    active = [list(range(n)) for n in n_tower] if info is None else info.active_idx
    slots = [None] + [slot_maps(active[l - 1], n_tower[l - 1], dev, model._slot_cache) for l in range(1, n_level)]
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (training and model.dropout_p > 0) else 0
    cfg = {"model": model, "info": info, "precise": model.expert_precision == "bf16x3", "seed": seed, "slots": slots,
           "want_gate_means": want_gate_means, "want_gates": want_gates, "want_gate_inputs": want_gate_inputs,
           "graph_after": GRAPH_AFTER if may_record else GRAPH_AFTER_CANDIDATE}
    if info is not None:
        info.warm(dev, n_level)            # device copies of the mask tables exist before any capture starts
    if STEP_SEED_PTR is not None and seed != 0:
        cfg["seed"], cfg["seed_ptr"] = 0, STEP_SEED_PTR
    if USE_GRAPHS and _mem.ENABLED and not want_gate_means and not want_gates:
        key = (0 if info is None else info.serial, tuple(x.shape), training, model.expert_precision,
               model.dropout_p if training else 0.0, torch.is_grad_enabled(), dev)
        entry = model._graphs.get(key)
        # record_graphs: one eager pass (sizes the arena, loads every kernel), the next one records
        entry.calls = entry.calls + 1 if not model._graphs.force else max(entry.calls + 1, cfg["graph_after"])
        cfg["graph"] = entry
    probs = AreadNode.apply(x, cfg, *model._fused_params)
    model.embedding.plan(dev).post_lookup(embedding_ops_bounds_mode())
    return probs, cfg


def embedding_ops_bounds_mode():
    from . import layer
    return layer.BOUNDS_MODE
