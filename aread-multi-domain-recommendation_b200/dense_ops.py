"""Host orchestration of the dense part of the AREAD step (trunk, HEI levels, heads, regulariser).

Reference: model/aread.py:131-153 (trunk), :263-322 (HEI under a HEMP mask), :156-202 (unmasked
walk), model/layer.py:96-112 (regulariser).
"""
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch


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

    def warm(self, device, n_level):
        """Create every cached device tensor now (they are host->device copies, which a stream capture refuses)."""
        self.group_index(device)
        for l in range(n_level):
            self.index(l, device)
            if l > 0:
                self.edges(l, device)

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


def aread_forward(model, x, info: Optional[MaskInfo], want_gate_means=False, want_gates=False,
                  want_gate_inputs=False, may_record=True) -> ForwardOut:
    """Embedding lookup -> trunk -> HEI levels -> per-tower probabilities, as one fused autograd node
    (fused.py)."""
    _require_cuda(model, x)
    from . import fused
    probs, cfg = fused.forward(model, x, info, want_gate_means, want_gates, want_gate_inputs, may_record)
    return ForwardOut(probs=probs, gate_inputs=cfg["gate_inputs"], gate_means=cfg["gate_means"], gates=cfg["gates"])


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
