"""torch-fp32 functional restatement of the AREAD hot path (TEST INFRASTRUCTURE, see
oracle/__init__.py).  Nothing here is an nn.Module: the model is a flat dict of tensors keyed
exactly like the reference `state_dict` plus a `Spec`.

Follows:
  * trunk                      /root/reference/model/aread.py:131-153
  * mode heads                 aread.py:156-202 (wo_mask), 224-234 (domain_with_mask),
                               235-244 (domain_mask_bagging)
  * HEI under a HEMP mask      aread.py:263-322
  * MLP / BN / cross / linear  /root/reference/model/layer.py:221-229, 529-537, 122-126
  * L2 regulariser             layer.py:96-112 with the groups registered at layer.py:31-33 and
                               aread.py:102-104, 123-127
  * loss + optimiser           /root/reference/run.py:672-682, 830-833
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .embedding_np import field_offsets

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass
class Spec:
    one_hot_field_dims: Sequence[int]
    embed_dim: int = 32
    multi_hot_flag: Sequence[bool] = ()
    itemid_idx: int = 0
    seq_maxlen: int = 5
    method: Optional[str] = "mean"
    n_tower: Sequence[int] = (3, 6, 12)
    n_domain: int = 30
    expert_dims: Sequence[int] = (256, 128, 64)
    tower_dims: Sequence[Sequence[int]] = ((64, 32), (32, 16), (16, 8))
    domain_idx: int = 10
    n_cross_layers: int = 3
    n_expert: int = 4
    dropout: float = 0.0
    l2_embedding: float = 1e-5
    l2_linear: float = 1e-5
    l2_dnn: float = 1e-5
    l2_cross: float = 1e-5
    # "bf16 experts" (BASELINE.json configs[1]): inputs and weights of the expert Linear layers are
    # rounded to bf16 (round to nearest even), products and sums stay fp32.  None = the reference's fp32.
    expert_operand_dtype: Optional[torch.dtype] = None
    # The CUDA path stores the bias-free pre-activation of every expert Linear ONCE, as bf16, whenever a backward may
    # follow (csrc/gemm_tc.cu STATS epilogue); BatchNorm statistics still come from the fp32 accumulator.  Setting
    # this reproduces that storage rounding (straight-through for the gradient); only meaningful together with
    # expert_operand_dtype.  None = no rounding (the reference's fp32, and the CUDA inference path).
    expert_preact_dtype: Optional[torch.dtype] = None
    flag: np.ndarray = field(init=False)

    def __post_init__(self):
        n_oh = len(self.one_hot_field_dims)
        fl = np.asarray(self.multi_hot_flag, dtype=bool)
        if fl.size == 0:
            fl = np.zeros(n_oh, dtype=bool)
        self.flag = fl
        self.n_mh_cols = int(fl.sum())
        self.n_mh_fields = self.n_mh_cols // self.seq_maxlen if self.n_mh_cols else 0
        pooled = self.method in ("mean", "sum")
        self.out_fields = n_oh + (self.n_mh_fields if pooled else self.n_mh_cols)
        self.E = self.out_fields * self.embed_dim
        self.n_rows = int(np.sum(self.one_hot_field_dims))
        self.offsets = field_offsets(self.one_hot_field_dims, fl, self.itemid_idx)
        self.n_level = len(self.n_tower)


# ----------------------------------------------------------------------------- building blocks
def embed(sd, spec, x):
    """layer.py:165-178 in torch (used for the float part; the bit-exact check lives in
    embedding_np.gather_fwd)."""
    idx = x + x.new_tensor(spec.offsets).unsqueeze(0)
    e = F.embedding(idx, sd["embedding.embedding_dict.weight"])
    if spec.n_mh_cols and spec.method in ("mean", "sum"):
        fl = torch.from_numpy(spec.flag).to(e.device)
        oh = e[:, ~fl, :]
        mh = e[:, fl, :].view(e.shape[0], spec.n_mh_fields, spec.seq_maxlen, spec.embed_dim)
        pooled = mh.mean(dim=2) if spec.method == "mean" else mh.sum(dim=2)
        e = torch.cat([oh, pooled], dim=1)
    return e


def _drop(h, p, training, masks, key):
    if not training or p <= 0.0:
        return h
    if isinstance(masks, str) and masks == "rng":      # timing runs: torch's own dropout stream
        return F.dropout(h, p, True)
    if masks is None:
        raise ValueError("oracle train-mode dropout needs explicit keep masks (dropout RNG is "
                         "implementation specific); pass dropout=0 for parity runs")
    keep = masks[key].to(h.dtype)
    return h * keep / (1.0 - p)


def mlp(sd, prefix, h, n_layers, training, p=0.0, masks=None, update_stats=True, operand_dtype=None,
        preact_dtype=None):
    """(Linear -> BatchNorm1d -> ReLU -> Dropout) x n; BN skipped when the batch is one row
    (layer.py:209-215, 225-228).  Running statistics in `sd` are updated in place in training."""
    for i in range(n_layers):
        li = 4 * i
        w = sd[f"{prefix}.layers.{li}.weight"]
        if operand_dtype is not None:      # rounding is not differentiated through (straight-through)
            h = h + (h.to(operand_dtype).to(h.dtype) - h).detach()
            w = w + (w.to(operand_dtype).to(w.dtype) - w).detach()
        if preact_dtype is not None and h.shape[0] != 1 and training:
            # what the kernels normalise: the rounded bias-free accumulator, with the statistics of the unrounded one
            bias, bn = sd[f"{prefix}.layers.{li}.bias"], f"{prefix}.layers.{li + 1}"
            acc = F.linear(h, w)
            acc_r = acc + (acc.to(preact_dtype).to(acc.dtype) - acc).detach()
            mean, var = acc.mean(dim=0), acc.var(dim=0, unbiased=False)
            h = (acc_r - mean) * torch.rsqrt(var + BN_EPS) * sd[bn + ".weight"] + sd[bn + ".bias"]
            if update_stats:
                m = acc.shape[0]
                with torch.no_grad():
                    sd[bn + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (mean + bias))
                    sd[bn + ".running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * (m / max(m - 1, 1)))
                sd[bn + ".num_batches_tracked"] += 1
            h = torch.relu(h)
            h = _drop(h, p, training, masks, f"{prefix}.{i}")
            continue
        h = F.linear(h, w, sd[f"{prefix}.layers.{li}.bias"])
        if h.shape[0] != 1:
            bn = f"{prefix}.layers.{li + 1}"
            track = training and update_stats
            # batch statistics (biased variance) in training, running statistics in eval; the
            # running update uses the unbiased variance and momentum 0.1 (torch BatchNorm1d)
            h = F.batch_norm(h,
                             sd[bn + ".running_mean"] if (track or not training) else None,
                             sd[bn + ".running_var"] if (track or not training) else None,
                             sd[bn + ".weight"], sd[bn + ".bias"], training, BN_MOMENTUM, BN_EPS)
            if track:
                sd[bn + ".num_batches_tracked"] += 1
        h = torch.relu(h)
        h = _drop(h, p, training, masks, f"{prefix}.{i}")
    return h


def cross(sd, spec, x0):
    """x_{k+1} = x0 * (w_k . x_k) + b_k + x_k (layer.py:533-537)."""
    c = x0
    for k in range(spec.n_cross_layers):
        s = c @ sd[f"cn.w.{k}.weight"].t()
        c = x0 * s + sd[f"cn.b.{k}"] + c
    return c


def trunk(sd, spec, x, training, masks=None, update_stats=True):
    """aread.py:131-153 without the attention branch, whose result is never read."""
    e = embed(sd, spec, x)
    dom = e[:, spec.domain_idx, :]
    X = e.flatten(start_dim=1)
    lin = X @ sd["linear.fc.weight"].t() + sd["linear.fc.bias"]
    cn = cross(sd, spec, X)
    hs = [mlp(sd, f"mmoe_experts.{k}", X, len(spec.expert_dims), training, spec.dropout, masks, update_stats,
              spec.expert_operand_dtype, spec.expert_preact_dtype)
          for k in range(spec.n_expert)]
    H = torch.stack(hs, dim=1)                                        # [B, n_expert, h]
    t0 = []
    for g in range(spec.n_tower[0]):
        a = torch.softmax(X @ sd[f"mmoe_gates.{g}.0.weight"].t() + sd[f"mmoe_gates.{g}.0.bias"], dim=1)
        t0.append((a.unsqueeze(-1) * H).sum(dim=1))
    return dict(X=X, dom=dom, lin=lin, cn=cn, t0=t0)


def _tower(sd, spec, l, t, h, training, masks, update_stats):
    return mlp(sd, f"towers.{l}.{t}", h, len(spec.tower_dims[l]), training, spec.dropout, masks, update_stats)


def hei(sd, spec, tr, q, mask, training, masks=None, update_stats=True):
    """Hierarchical expert integration, optionally under a HEMP mask (aread.py:263-322;
    mask=None is the unmasked walk of aread.py:164-186).

    Returns (probs, active_last, gate_means, gates):
      probs        list over the *active* last-level towers of [B] probabilities
      gate_means   {(l, t): mean_b(softmax * mask column)}   (zeros for inactive towers)
      gates        {(l, t): [B, n_{l-1}] raw softmax}        (the wo_mask side output)"""
    B = tr["X"].shape[0]
    nl = spec.n_level
    gate_means, gates = {}, {}
    prev = None
    probs, active_last = [], []
    for l in range(nl):
        n_t = spec.n_tower[l]
        act = [True] * n_t if mask is None else [bool(mask[l][:, t].any()) for t in range(n_t)]
        outs = []
        for t in range(n_t):
            width = spec.tower_dims[l][-1]
            if not act[t]:
                outs.append(torch.zeros(B, width, device=q.device))
                if l > 0:
                    gate_means[(l, t)] = torch.zeros(spec.n_tower[l - 1], device=q.device)
                continue
            if l == 0:
                inp = tr["t0"][t]
            else:
                s = torch.softmax(q @ sd[f"tower_gates.{l - 1}.{t}.0.weight"].t()
                                  + sd[f"tower_gates.{l - 1}.{t}.0.bias"], dim=1)
                gates[(l, t)] = s.detach()
                if mask is None:
                    r = s
                else:
                    col = mask[l][:, t].to(s.dtype)
                    sm = s * col
                    r = sm / (sm.sum(dim=1, keepdim=True) + 1e-8)
                    gate_means[(l, t)] = sm.mean(dim=0).detach()
                inp = (r.unsqueeze(-1) * prev).sum(dim=1)
            outs.append(_tower(sd, spec, l, t, inp, training, masks, update_stats))
        if l < nl - 1:
            prev = torch.stack(outs, dim=1)
        else:
            for t in range(n_t):
                if act[t]:
                    z = torch.cat([tr["cn"], outs[t]], dim=1) @ sd[f"towers_linear.{t}.weight"].t() + tr["lin"]
                    probs.append(torch.sigmoid(z).squeeze(-1))
                    active_last.append(t)
    return probs, active_last, gate_means, gates


def forward(sd, spec, x, mode, mask=None, training=False, masks=None, update_stats=True):
    """The three live modes.  Returns a dict with `y` (what the reference forward returns),
    `gate_means` and `gates`."""
    tr = trunk(sd, spec, x, training, masks, update_stats)
    if mode == "wo_mask":
        q = torch.cat([tr["dom"], torch.zeros_like(tr["dom"])], dim=1)
        probs, _, gm, gates = hei(sd, spec, tr, q, None, training, masks, update_stats)
        y = torch.stack(probs, dim=0).unsqueeze(-1).mean(dim=0)           # [B, 1]
        return dict(y=y, gate_means=gm, gates=gates)
    if mode not in ("domain_with_mask", "domain_mask_bagging"):
        raise ValueError(mode)
    act0 = torch.nonzero(mask[0])[:, 1]
    grp = sd["group_embedding.weight"][act0]
    if grp.shape[0] > 1:
        grp = grp.mean(dim=0, keepdim=True)
    q = torch.cat([tr["dom"], grp.expand(x.shape[0], -1)], dim=1)
    probs, _, gm, gates = hei(sd, spec, tr, q, mask, training, masks, update_stats)
    ys = torch.stack(probs, dim=0)
    y = ys if mode == "domain_mask_bagging" else ys.mean(dim=0)
    return dict(y=y, gate_means=gm, gates=gates)


# ----------------------------------------------------------------------------- loss / step
def reg_groups(sd, spec):
    """[(l2, [keys])] in registration order (layer.py:31-33, aread.py:102-104, 123-127).  The
    name filter is `'weight' in name and 'bn' not in name`; BN layers are called `layers.N`
    so their gamma passes it."""
    groups = [(spec.l2_embedding, ["embedding.embedding_dict.weight"]),
              (spec.l2_linear, ["linear.fc.weight"])]
    groups.append((spec.l2_dnn, [k for k in sd if k.startswith("mmoe_experts.") and k.endswith(".weight")]))
    groups.append((spec.l2_dnn, [k for k in sd if k.startswith("towers.") and k.endswith(".weight")]))
    groups.append((spec.l2_cross, [k for k in sd if k.startswith("cn.") and k.endswith(".weight")]))
    return groups


def reg_loss(sd, spec):
    total = torch.zeros(1, device=sd["linear.fc.weight"].device)
    for l2, keys in reg_groups(sd, spec):
        for k in keys:
            total = total + torch.sum(l2 * torch.square(sd[k]))
    return total


def bagging_loss(y_stack, target):
    """sum_t BCE_mean(y_stack[t], y) / n_active (run.py:672-677)."""
    target = target.reshape(-1).float()
    losses = [F.binary_cross_entropy(p, target) for p in y_stack.unbind(dim=0)]
    return sum(losses) / y_stack.shape[0]


PARAM_SUFFIXES = (".weight", ".bias")


def is_param(key, sd=None):
    if key.endswith(("running_mean", "running_var", "num_batches_tracked")):
        return False
    return True


def make_leaf_params(sd):
    """Turn every trainable entry of `sd` into a leaf that requires grad (in place)."""
    for k in list(sd):
        if is_param(k) and sd[k].is_floating_point():
            sd[k] = sd[k].detach().clone().requires_grad_(True)
    return sd


def train_step(sd, spec, x, target, mask, opt, mode="domain_mask_bagging", masks=None):
    """forward -> loss (+ L2) -> zero_grad(None) -> backward -> Adam step.  `opt` is a
    torch.optim.Adam over the leaves of `sd` built with the trainer's settings
    (run.py:830-831: betas (0.9, 0.99), eps 1e-8, coupled weight_decay 1e-8)."""
    out = forward(sd, spec, x, mode, mask, training=True, masks=masks)
    if mode == "wo_mask":
        loss = F.binary_cross_entropy(out["y"].squeeze(), target.reshape(-1).float())
    else:
        loss = bagging_loss(out["y"], target)
    loss = loss + reg_loss(sd, spec)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss.detach(), out


def make_adam(sd, lr=1e-3, wd=1e-8):
    leaves = [v for v in sd.values() if isinstance(v, torch.Tensor) and v.requires_grad]
    return torch.optim.Adam(leaves, lr=lr, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
