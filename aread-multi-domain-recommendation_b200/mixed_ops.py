"""Mixed-domain batches (BASELINE north_star: domain-sorted, per-(domain, tower) tile skipping): the eval-mode forward
of `mode='domain_with_mask'` for rows of MANY domains in one call, each row under `domain_mask[its domain]`, and the
per-domain gate means of a mixed `wo_mask` batch.

Eval mode is row-local (BatchNorm on running statistics, no dropout, gates from the row's own domain embedding and
its domain's mask), so the result equals the reference called once per domain (run.py:719-727,
model/aread.py:224-234, 263-322) row for row; tests/test_mixed_gpu.py checks it against the per-domain calls and the
oracle.  Trunk = the usual kernels (lookup, row pass on the tensor cores, bf16 experts with BatchNorm + ReLU folded into
the GEMM epilogue, MMoE mixture) with every tower's gate / head evaluated; HEI levels = csrc/mixed.cu.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import dense_kernels as dk
from . import embedding_ops
from . import rowpass_ops
from . import tower_ops
from .expert_ops import _hi, _lo


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class MaskTables:
    """Per-domain bit tables of the installed masks (`domain_mask`), kept on the device until a mask changes."""

    def __init__(self, model, device):
        n_level, n_tower, n_domain = model.n_level, model.n_tower, model.n_domain
        infos = [model.mask_info(model.domain_mask[d]) for d in range(n_domain)]
        self.key = tuple(i.serial for i in infos)
        active = np.zeros((n_domain, n_level), dtype=np.uint32)
        edge_words = int(sum(n_tower[1:]))
        edges = np.zeros((n_domain, max(edge_words, 1)), dtype=np.uint32)
        group = np.zeros((n_domain, n_tower[0]), dtype=np.float32)
        for d, info in enumerate(infos):
            base = 0
            for l in range(n_level):
                for t in info.active_idx[l]:
                    active[d, l] |= np.uint32(1 << t)
                if l > 0:
                    arr = info.arrays[l]                                   # [n_{l-1}, n_l]
                    for t in range(n_tower[l]):
                        bits = 0
                        for j in np.nonzero(arr[:, t])[0]:
                            bits |= 1 << int(j)
                        edges[d, base + t] = np.uint32(bits)
                    base += n_tower[l]
            g = info.group_idx                                             # aread.py:226: nonzero(mask[0])
            if len(g):
                np.add.at(group[d], g, 1.0 / len(g))
        self.edge_words = edge_words
        self.active = torch.from_numpy(active.view(np.int32)).to(device)
        self.edges = torch.from_numpy(edges.view(np.int32)).to(device)
        self.group_weights = torch.from_numpy(group).to(device)             # [n_domain, n0]: grp = weights @ G_emb


def mask_tables(model, device):
    key = tuple(model.mask_info(model.domain_mask[d]).serial for d in range(model.n_domain))
    cached = model._mixed_tables.get(device)
    if cached is None or cached.key != key:
        cached = model._mixed_tables[device] = MaskTables(model, device)
    return cached


def _ptr_grid(lib_field, values):
    for l, row in enumerate(values):
        for j, t in enumerate(row):
            lib_field[l][j] = t.data_ptr() if t is not None else None


@torch.no_grad()
def forward_mixed_eval(model, x, want_stack=False, hei_events=None):
    """y [B] (and y_stack [n_last, B] with want_stack): the 'domain_with_mask' probabilities of a mixed batch in
    eval mode; row b uses domain_mask[x[b, domain_idx]]."""
    if model.training:
        raise RuntimeError("forward_mixed is the eval-mode path (BatchNorm on running statistics); call model.eval()")
    if not x.is_cuda:
        raise RuntimeError("aread_b200: inputs must be CUDA tensors -- no CPU fallback")
    lib = _lib.load()
    P = model._fused
    table = model.embedding.embedding_dict.weight
    x = embedding_ops.prepare_ids(x, table)
    dev = x.device
    B = x.shape[0]
    D, E = model.embed_dim, model.embed_output_dim
    n_level, n_tower = model.n_level, model.n_tower
    n_expert, n_cross = len(model.mmoe_experts), model.cn.num_layers
    bn_skip = B == 1
    precise = model.expert_precision == "bf16x3"
    tabs = mask_tables(model, dev)

    # ---- lookup + row pass with EVERY gate and head (the mask picks per row later)
    plan = model.embedding.plan(dev)
    embed, (x_hi, x_lo) = embedding_ops.gather(plan, table, x, want_bf16=True, want_lo=True, fence=True)
    X = embed.view(B, E)
    w_out = P.tl_w.flat[:, 0, :]
    w_cat = torch.cat([model.linear.fc.weight, P.mmoe_w.flat.reshape(-1, E), P.cn_w.flat.view(n_cross, E), w_out[:, :E]], dim=0)
    beta = torch.cumsum(P.cn_b.flat, dim=0)
    kappa = torch.zeros(n_cross, dtype=torch.float32, device=dev)
    if n_cross > 1:
        kappa[1:] = (P.cn_w.flat.view(n_cross, E)[1:] * beta[:-1]).sum(dim=1)
    beta_n = beta[-1] if n_cross > 0 else torch.zeros(E, dtype=torch.float32, device=dev)
    offset = torch.cat([model.linear.fc.bias, P.mmoe_b.flat.reshape(-1), kappa, w_out[:, :E] @ beta_n], dim=0)
    layout = (n_tower[0], n_expert, n_cross, n_tower[-1])
    nj = w_cat.shape[0]
    ldp = (nj + 3) // 4 * 4
    p_dots = torch.empty((B, ldp), dtype=torch.float32, device=dev)
    lin = torch.empty((B,), dtype=torch.float32, device=dev)
    gate = torch.empty((B, n_tower[0], n_expert), dtype=torch.float32, device=dev)
    alpha = torch.empty((B, n_cross + 1), dtype=torch.float32, device=dev)
    head_cross = torch.empty((B, n_tower[-1]), dtype=torch.float32, device=dev)
    from . import fused
    tc_row = fused.TC_ROWPASS and not precise and nj <= 128 and E % 8 == 0       # as in the fused train / eval node
    if tc_row:
        wc_hi, wc_lo = dk.split_bf16(w_cat)
        dk.grouped_linear(x_hi, wc_hi, None, nj, E, 1, 0, out=p_dots, a_lo=x_lo, w_lo=wc_lo, lo_lo=True)
    ra = rowpass_ops._args(B, E, layout, ldp, x=None if tc_row else X, w=None if tc_row else w_cat, offset=offset,
                           p=p_dots, lin=lin, gate=gate, alpha=alpha, head=head_cross)
    _lib.check(lib.aread_rowpass_fwd(ctypes.byref(ra), _stream(dev)))

    # ---- experts + MMoE mixture for all level-0 towers
    G = P.experts[0].groups
    a_op = (x_hi, x_lo) if precise else x_hi
    z = stats = None
    if precise or not all(L.n % 64 == 0 for L in P.experts):      # (narrow expert layers: unfused kernels)
        for i, L in enumerate(P.experts):
            w = dk.split_bf16(L.weight.rows()) if precise else L.weight.rows().to(torch.bfloat16)
            z = dk.grouped_linear(_hi(a_op), _hi(w), L.bias.rows(), L.n, L.k, G, 0 if i == 0 else L.k, a_lo=_lo(a_op),
                                  w_lo=_lo(w))
            last = i == len(P.experts) - 1
            res = dk.bn_act_fwd(z, L.gamma.rows(), L.beta.rows(), L.running_mean.rows(), L.running_var.rows(), False,
                                bn_skip, 0.0, 0, L.salt, None if last else torch.bfloat16, want_lo=precise and not last)
            a_op, stats = (None, res[-1]) if last else (((res[0], res[1]), res[2]) if precise else (res[0], res[1]))
    else:
        w16 = model.expert_weights_bf16()
        for i, L in enumerate(P.experts):
            folded = dk.expert_bn_finalize(None, B, G * L.n, L.bias.rows(), L.gamma.rows(), L.beta.rows(),
                                           L.running_mean.rows(), L.running_var.rows(), False, bn_skip)
            a_op = z = dk.expert_linear_act(a_op, w16[i], L.n, L.k, G, 0 if i == 0 else L.k, folded)
        stats = dk.identity_saved(G * P.experts[-1].n, dev)
    t0 = dk.mmoe_mix_fwd(z, stats, gate, G, n_tower[0], 0.0, 0, P.experts[-1].salt)            # [B, n0, H]

    # ---- gate logits of every tower from q = [domain embedding | mean group embedding of the row's domain]
    dom_ids = x[:, model.domain_idx].long().clamp_(0, model.n_domain - 1)
    grp = tabs.group_weights @ model.group_embedding.weight                                    # [n_domain, D]
    q = torch.cat([embed[:, model.domain_idx, :], grp.index_select(0, dom_ids)], dim=1)        # [B, 2D]
    logits = [None] * n_level
    for l in range(1, n_level):
        n_prev = n_tower[l - 1]
        wg, bg = P.gate_w[l].flat, P.gate_b[l].flat
        if n_tower[l] * n_prev <= 128:
            logits[l] = tower_ops.tower_linear(q, wg.view(1, n_tower[l] * n_prev, -1), bg.reshape(-1), n_tower[l] * n_prev,
                                               groups=1)
        else:
            logits[l] = tower_ops.tower_linear(q, wg, bg.reshape(-1), n_prev, groups=n_tower[l])

    # ---- HEI levels + heads, every row under its own domain's mask
    a = _lib.HeiMixedArgs()
    a.m, a.n_level, a.n_layer = B, n_level, len(P.towers[0])
    for l in range(n_level):
        a.n_tower[l] = n_tower[l]
        for j, L in enumerate(P.towers[l]):
            a.dims[l][j] = L.n
            a.weight[l][j] = L.weight.flat.data_ptr()
            a.bias[l][j] = L.bias.flat.data_ptr()
            a.gamma[l][j] = L.gamma.flat.data_ptr()
            a.beta[l][j] = L.beta.flat.data_ptr()
            a.running_mean[l][j] = L.running_mean.flat.data_ptr()
            a.running_var[l][j] = L.running_var.flat.data_ptr()
        if l > 0:
            a.logits[l] = logits[l].data_ptr()
    a.width_in = P.towers[0][0].k
    a.n_domain, a.edge_words, a.bn_skip, a.eps = model.n_domain, tabs.edge_words, 1 if bn_skip else 0, dk.BN_EPS
    a.domain, a.domain_stride = x.data_ptr() + 4 * model.domain_idx, x.stride(0)
    a.active, a.edges = tabs.active.data_ptr(), tabs.edges.data_ptr()
    a.t0, a.head_cross, a.lin = t0.data_ptr(), head_cross.data_ptr(), lin.data_ptr()
    w_tail = w_out[:, E:].contiguous()
    a.w_tail = w_tail.data_ptr()
    y = torch.empty((B,), dtype=torch.float32, device=dev)
    y_stack = torch.empty((n_tower[-1], B), dtype=torch.float32, device=dev) if want_stack else None
    a.y, a.y_stack = y.data_ptr(), y_stack.data_ptr() if want_stack else None
    if not lib.aread_hei_mixed_eval_supported(ctypes.byref(a)):
        raise RuntimeError("forward_mixed: this tower configuration does not fit the shared memory of one SM")
    if hei_events is not None:          # (bench.py: the HEI kernel timed on its own, on its launch stream)
        hei_events[0].record(torch.cuda.current_stream(dev))
    _lib.check(lib.aread_hei_mixed_eval(ctypes.byref(a), _stream(dev)))
    if hei_events is not None:
        hei_events[1].record(torch.cuda.current_stream(dev))
    model.embedding.plan(dev).post_lookup(_bounds_mode())
    return (y, y_stack) if want_stack else y


def _bounds_mode():
    from . import layer
    return layer.BOUNDS_MODE


def domain_means(values, x, domain_idx, n_domain):
    """(mean [n_domain, C], count [n_domain]) of values [B, C] over the rows of each domain (batch order)."""
    B, C = values.shape
    values = values.contiguous()
    mean = torch.empty((n_domain, C), dtype=torch.float32, device=values.device)
    count = torch.empty((n_domain,), dtype=torch.int32, device=values.device)
    a = _lib.DomainMeanArgs(B, C, n_domain, values.data_ptr(), values.stride(0), x.data_ptr() + 4 * domain_idx, x.stride(0),
                            mean.data_ptr(), count.data_ptr())
    _lib.check(_lib.load().aread_domain_mean(ctypes.byref(a), _stream(values.device)))
    return mean, count
