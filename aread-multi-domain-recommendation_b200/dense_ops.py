"""Host orchestration of the dense part of the AREAD step (trunk, HEI levels, heads, regulariser).

Reference: model/aread.py:131-153 (trunk), :263-322 (HEI under a HEMP mask), :156-202 (unmasked
walk), model/layer.py:96-112 (regulariser).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import embedding_ops, expert_ops, rowpass_ops, tower_ops


class MaskInfo:
    """Host view of one HEMP mask: which towers run, which edges are open.  Built once per distinct
    mask; the forward consults it instead of synchronising on device booleans."""
    _serials = 0

    def __init__(self, arrays, n_tower):
        MaskInfo._serials += 1
        self.serial = MaskInfo._serials                              # never reused (keys the CUDA-graph cache)
        self.arrays = [np.asarray(a, dtype=bool) for a in arrays]
        self.n_tower = tuple(n_tower)
        n_level = len(self.n_tower)
        # a tower runs when any edge enters it (aread.py:268: any(mask[l], dim=0))
        self.active = [self.arrays[l].any(axis=0) for l in range(n_level)]
        self.active_last = np.nonzero(self.active[-1])[0]
        self.active_idx = [[int(t) for t in np.nonzero(a)[0]] for a in self.active]
        self.group_idx = np.nonzero(self.arrays[0])[1]            # aread.py:226 / 237
        self._edges = {}

    def edges(self, l, device):
        key = (l, device)
        e = self._edges.get(key)
        if e is None:
            e = torch.from_numpy(self.arrays[l].astype(np.float32)).to(device)
            self._edges[key] = e
        return e

    def index(self, l, device):
        """int64 device tensor of the towers of level l that run."""
        key = ("idx", l, device)
        e = self._edges.get(key)
        if e is None:
            e = torch.tensor(self.active_idx[l], dtype=torch.int64, device=device)
            self._edges[key] = e
        return e

    def group_index(self, device):
        key = ("grp", device)
        e = self._edges.get(key)
        if e is None:
            e = torch.from_numpy(self.group_idx.astype(np.int64)).to(device)
            self._edges[key] = e
        return e


@dataclass
class ForwardOut:
    probs: torch.Tensor                                  # [n_active_last, B]
    gate_inputs: Optional[torch.Tensor] = None           # [B, 2 * D]
    gate_means: Dict[int, torch.Tensor] = field(default_factory=dict)   # l -> [n_{l-1}, n_l]
    gates: Dict[int, torch.Tensor] = field(default_factory=dict)        # l -> [B, n_{l-1}, n_l]


def _require_cuda(model, x):
    if not x.is_cuda:
        raise RuntimeError("aread_b200: inputs must be CUDA tensors -- this implementation is sm_100a-only and "
                           "has no CPU fallback")


def aread_forward(model, x, info: Optional[MaskInfo], want_gate_means=False, want_gates=False) -> ForwardOut:
    """Embedding lookup -> trunk -> HEI levels -> per-tower probabilities, as one fused autograd node
    (fused.py)."""
    _require_cuda(model, x)
    from . import fused
    probs, cfg = fused.forward(model, x, info, want_gate_means, want_gates)
    return ForwardOut(probs=probs, gate_inputs=cfg["gate_inputs"], gate_means=cfg["gate_means"], gates=cfg["gates"])


def hei_levels(model, level0_in, q, head_cross, lin, info: Optional[MaskInfo], seed, want_gate_means=False,
               want_gates=False) -> ForwardOut:
    """The HEI levels on compact activations ([B, n_active, width]: only towers that run).

    level0_in [B, n_active0, h]: MMoE mixtures of the active level-0 towers; head_cross [B, n_active_last]:
    the cross-network part of the active heads; lin [B]."""
    n_level, n_tower = model.n_level, model.n_tower
    dev = q.device
    training, p = model.training, model.dropout_p
    out = ForwardOut(probs=None)
    h = level0_in
    prev_active = None
    for l in range(n_level):
        active = list(range(n_tower[l])) if info is None else info.active_idx[l]
        index = None if (info is None or len(active) == n_tower[l]) else info.index(l, dev)
        if l > 0:
            # gates over ALL towers of the previous level (aread.py:282-285), evaluated only for towers that run
            w_g = torch.stack([model.tower_gates[l - 1][t][0].weight for t in active], dim=0)   # [na, n_prev, 2D]
            b_g = torch.stack([model.tower_gates[l - 1][t][0].bias for t in active], dim=0)     # [na, n_prev]
            s = torch.softmax(torch.einsum('bd,tjd->btj', q, w_g) + b_g, dim=2)                 # [B, na, n_prev]
            if want_gates:
                out.gates[l] = s.detach().transpose(1, 2)                                       # [B, n_prev, n_l]
            if info is None:
                r = s
            else:
                edges = info.edges(l, dev).t()                                                  # [n_l, n_prev]
                sm = s * (edges if index is None else edges.index_select(0, index))
                r = sm / (sm.sum(dim=2, keepdim=True) + 1e-8)
                if want_gate_means:
                    means = sm.detach().mean(dim=0).t()                                         # [n_prev, na]
                    if index is not None:       # towers that do not run report zeros (aread.py:278-280)
                        means = torch.zeros(n_tower[l - 1], n_tower[l], dtype=torch.float32,
                                            device=dev).index_copy_(1, index, means)
                    out.gate_means[l] = means
            if len(prev_active) != n_tower[l - 1]:
                r = r.index_select(2, info.index(l - 1, dev))
            h = torch.einsum('btj,bjw->btw', r, h)                                              # [B, na, w_prev]
        for layer in model._tower_layers[l]:
            h = tower_ops.tower_layer(h, layer, active, index, training, p, seed)
        prev_active = active
    E = model.embed_output_dim
    w_tail = torch.stack([model.towers_linear[t].weight[0, E:] for t in prev_active], dim=0)     # [na, w_last]
    z = head_cross + (h * w_tail).sum(dim=2) + lin.unsqueeze(1)                                  # [B, na]
    out.probs = torch.sigmoid(z).t()
    return out


def hei_forward(model, tower_inputs, q, cn, lin, info: Optional[MaskInfo], want_gate_means=False,
                want_gates=False, head_cross=None) -> ForwardOut:
    """`head_cross[t]` ([B, 1]) is the cross-network part of head t (w_out_t[:E] . cn_out) from the row
    pass; without it the heads are evaluated from an explicit `cn` tensor like the reference does."""
    B = lin.shape[0]
    n_level, n_tower = model.n_level, model.n_tower
    out = ForwardOut(probs=None)
    prev = None
    for l in range(n_level):
        active = np.ones(n_tower[l], dtype=bool) if info is None else info.active[l]
        width = model.tower_dims[l][-1]
        if l > 0:
            # gates of towers that do not run are never evaluated (their parameters keep grad None)
            zeros = torch.zeros(B, n_tower[l - 1], dtype=torch.float32, device=q.device)
            logits = torch.stack([model.tower_gates[l - 1][t][0](q) if active[t] else zeros
                                  for t in range(n_tower[l])], dim=2)
            s = torch.softmax(logits, dim=1)                                        # [B, n_{l-1}, n_l]
            if want_gates:
                out.gates[l] = s.detach()
            if info is None:
                r = s
            else:
                sm = s * info.edges(l, q.device)
                r = sm / (sm.sum(dim=1, keepdim=True) + 1e-8)
                if want_gate_means:
                    out.gate_means[l] = sm.mean(dim=0).detach()
            inputs = torch.einsum('bjt,bjw->btw', r, prev)                          # [B, n_l, w_{l-1}]
        level_out = []
        for t in range(n_tower[l]):
            if not active[t]:
                level_out.append(torch.zeros(B, width, dtype=torch.float32, device=q.device))
                continue
            h = tower_inputs[t] if l == 0 else inputs[:, t, :]
            level_out.append(model.towers[l][t](h))
        if l < n_level - 1:
            prev = torch.stack(level_out, dim=1)
        else:
            probs = []
            for t in range(n_tower[l]):
                if active[t]:
                    if head_cross is not None:
                        E = model.embed_output_dim
                        z = head_cross[t] + level_out[t] @ model.towers_linear[t].weight[:, E:].t() + lin
                    else:
                        z = model.towers_linear[t](torch.cat([cn, level_out[t]], dim=1)) + lin
                    probs.append(torch.sigmoid(z).squeeze(-1))
            out.probs = torch.stack(probs, dim=0)
    return out


def regularization_loss(regularization_weight, device):
    """sum_groups sum_w l1*|w| + l2*w^2 as a [1] tensor (layer.py:96-112)."""
    total = torch.zeros((1,), device=device)
    for weights, l1, l2 in regularization_weight:
        for w in weights:
            p = w[1] if isinstance(w, tuple) else w
            if l1 > 0:
                total = total + torch.sum(l1 * torch.abs(p))
            if l2 > 0:
                total = total + l2 * torch.sum(torch.square(p))
    return total
