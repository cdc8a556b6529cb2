"""Host-side mirror of the reference `model/layer.py` for the AREAD hot path.

Same class names, constructor / forward signatures, attribute names and `state_dict` keys as the
reference, so `run.py` and checkpoints keep working; the arithmetic runs in
libaread_sm100.so.  Classes mirrored here: BaseModel (layer.py:9-112), FeaturesLinear (:115-126),
FeaturesEmbedding (:129-183), MultiLayerPerceptron (:203-229), CrossNetwork (:517-537).  Every
other symbol of the reference module (baseline-only layers) is resolved lazily from the reference
tree, see `__getattr__` at the bottom.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
from torch import nn

from . import embedding_ops

BOUNDS_MODE = os.environ.get("AREAD_BOUNDS_CHECK", "deferred")  # 'sync' | 'deferred' | 'off'


class FeaturesEmbedding(nn.Module):
    """One shared table for all fields; ids are shifted by per-column row offsets, multi-hot
    columns re-use the item-id field's offset and are pooled over `seq_maxlen` positions."""

    def __init__(self, one_hot_field_dims, embed_dim, multi_hot_dict=None):
        super().__init__()
        multi_hot_dict = multi_hot_dict or {"multi_hot_flag": [False] * len(one_hot_field_dims), "itemid_idx": 0,
                                            "seq_maxlen": 1, "method": None}
        self.multi_hot_flag = np.array(multi_hot_dict["multi_hot_flag"])
        self.seq_maxlen = multi_hot_dict["seq_maxlen"]
        self.multi_hot_method = multi_hot_dict["method"]
        if self.multi_hot_method not in {"sum", "mean", None}:
            raise ValueError(f"Invalid multi-hot method '{self.multi_hot_method}'. "
                             "Method must be 'mean', 'sum', or None.")
        n_mh_cols = int(np.sum(self.multi_hot_flag))
        self.one_hot_field_num = len(one_hot_field_dims)
        self.multi_hot_field_num = n_mh_cols // self.seq_maxlen
        pooled = self.multi_hot_method in {"sum", "mean"}
        self.output_dim0 = self.one_hot_field_num + (self.multi_hot_field_num if pooled else n_mh_cols)
        self.embed_dim = embed_dim

        dims = np.asarray(one_hot_field_dims, dtype=np.int64)
        self._n_rows = int(dims.sum())
        self.embedding_dict = nn.Embedding(self._n_rows, embed_dim)         # N(0, 1) init, as the reference
        starts = np.concatenate(([0], np.cumsum(dims)[:-1])).astype(np.int64)
        if self.multi_hot_field_num > 0:
            starts = np.concatenate((starts, np.full(n_mh_cols, starts[multi_hot_dict["itemid_idx"]])))
        self.offsets = starts
        self._plans = {}

    def plan(self, device):
        key = (device.type, device.index)
        plan = self._plans.get(key)
        if plan is None:
            flag = self.multi_hot_flag
            if flag.size != len(self.offsets):      # flag shorter than x: treat the missing columns as one-hot
                flag = np.concatenate((flag, np.zeros(len(self.offsets) - flag.size, dtype=bool)))
            plan = embedding_ops.LookupPlan(self.offsets, flag, self.seq_maxlen, self.multi_hot_method,
                                            self.embed_dim, self._n_rows, device)
            self._plans[key] = plan
        return plan

    def shard_table(self, group=None):
        """Row-shard the table over the ranks of `group` (call after .to(device) and before building the
        optimizer): `embedding_dict.weight` becomes this rank's [ceil(R / world), D] shard, lookups read
        the other shards over NVLink and the gradient arrives reduce-scattered (sharding.py)."""
        import torch.distributed as dist
        from . import sharding
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        full = self.embedding_dict.weight.data
        n_rows = full.shape[0]
        self.embedding_dict.weight = nn.Parameter(sharding.split_table(full, world, rank))
        shards = sharding.TableShards(self.embedding_dict.weight, n_rows, group)
        plan = self.plan(full.device)
        plan.shards = shards
        return shards

    def lookup(self, x, want_bf16=False, want_lo=False):
        """(fp32 [batch, output_dim0, embed_dim], bf16 [batch, output_dim0 * embed_dim] or None); with
        want_lo the second element is the split pair (hi, lo)."""
        table = self.embedding_dict.weight
        x = embedding_ops.prepare_ids(x, table)
        plan = self.plan(x.device)
        res = embedding_ops.EmbeddingLookup.apply(table, x, plan, want_bf16, want_lo)
        plan.post_lookup(BOUNDS_MODE)
        if want_bf16 and want_lo:
            return res[0], (res[1], res[2])
        return res if want_bf16 else (res, None)

    def forward(self, x, squeeze_dim=False):
        """x: integer tensor (batch, n_cols) -> (batch, output_dim0, embed_dim) fp32."""
        out, _ = self.lookup(x)
        return out.flatten(start_dim=1) if squeeze_dim else out


class FeaturesLinear(nn.Module):
    """Linear term + bias over the flattened embedding."""

    def __init__(self, field_dims, output_dim=1, sigmoid=False):
        super().__init__()
        self.fc = nn.Linear(field_dims, output_dim, bias=True)
        self.sigmoid = sigmoid

    def forward(self, x):
        y = self.fc(x)
        return torch.sigmoid(y) if self.sigmoid else y


class MultiLayerPerceptron(nn.Module):
    """`layers` keeps the reference's slot numbering (Linear, BatchNorm1d, ReLU, Dropout per hidden
    layer) so that state_dict keys are `layers.{0,1,4,5,...}`."""

    def __init__(self, input_dim, layer_dims, dropout, output_layer=True, bn=True):
        super().__init__()
        self.layers = nn.ModuleList()
        width = input_dim
        for out_width in layer_dims:
            self.layers.append(nn.Linear(width, out_width))
            if bn:
                self.layers.append(nn.BatchNorm1d(out_width))
            self.layers.append(nn.ReLU())
            self.layers.append(nn.Dropout(p=dropout))
            width = out_width
        if output_layer:
            self.layers.append(nn.Linear(width, 1))

    def forward(self, x):
        single_row = x.shape[0] == 1       # BatchNorm is skipped for a batch of one (layer.py:226)
        for layer in self.layers:
            if single_row and isinstance(layer, nn.BatchNorm1d):
                continue
            x = layer(x)
        return x


class CrossNetwork(nn.Module):
    """DCN cross layers: x_{k+1} = x_0 * (w_k . x_k) + b_k + x_k."""

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.w = nn.ModuleList([nn.Linear(input_dim, 1, bias=False) for _ in range(num_layers)])
        self.b = nn.ParameterList([nn.Parameter(torch.zeros((input_dim,))) for _ in range(num_layers)])

    def forward(self, x):
        x0, xk = x, x
        for w, b in zip(self.w, self.b):
            xk = x0 * w(xk) + b + xk
        return xk


class BaseModel(nn.Module):
    """Embedding + linear term + the L1/L2 regulariser registry shared by all reference models."""

    def __init__(self, one_hot_feature_dims, embed_dim, multi_hot_dict, l2_reg_embedding=1e-5, l2_reg_linear=1e-5):
        super().__init__()
        self.multi_hot_flag = multi_hot_dict["multi_hot_flag"]
        self.feature_dims = one_hot_feature_dims + sum(self.multi_hot_flag)
        self.embedding = FeaturesEmbedding(one_hot_feature_dims, embed_dim, multi_hot_dict)
        self.embed_output_dim = self.embedding.output_dim0 * embed_dim
        self.embed_dim = embed_dim
        self.field_num = self.embedding.one_hot_field_num + self.embedding.multi_hot_field_num
        self.linear = FeaturesLinear(self.embed_output_dim)
        self.is_concat_linear_cn = None
        self.reg_loss = torch.zeros((1,))
        self.regularization_weight = []
        self.add_regularization_weight(self.embedding.embedding_dict.parameters(), l2=l2_reg_embedding)
        self.add_regularization_weight(_weights_without_bn(self.linear), l2=l2_reg_linear)

    # -- attention branch (state_dict compatibility; AREAD never reads its output, aread.py:139-140)
    def build_atten(self, config, dropout):
        width = getattr(config, "atten_embed_dim", self.embed_dim)
        self.atten_embedding = nn.Linear(self.embed_dim, width)
        self.atten_output_dim = self.embedding.output_dim0 * width
        self.att_res = config.att_res
        self.self_attns = nn.ModuleList(
            [nn.MultiheadAttention(config.atten_embed_dim, config.att_head_num, dropout=dropout)
             for _ in range(config.att_layer_num)])
        if self.att_res:
            self.V_res_embedding = nn.Linear(self.embed_dim, width)
        self.atten_linear = nn.Linear(self.atten_output_dim, 1, bias=False)

    def atten_forward(self, embed_x):
        fields = embed_x.reshape(-1, self.field_num, self.embed_dim)
        h = self.atten_embedding(fields).transpose(0, 1)           # (fields, batch, width)
        for attn in self.self_attns:
            h, _ = attn(h, h, h)
        h = h.transpose(0, 1)
        if self.att_res:
            h = h + self.V_res_embedding(fields)
        return self.atten_linear(torch.relu(h).reshape(-1, self.atten_output_dim))

    # -- multi-tower helpers used by the baseline models
    def build_tower_output(self, n_tower, tower_input_dim, tower_dims, dropout):
        towers = nn.ModuleList(MultiLayerPerceptron(tower_input_dim, tower_dims, dropout, output_layer=True)
                               for _ in range(self.n_tower))
        return towers, None, nn.ModuleList([nn.Sigmoid() for _ in range(n_tower)])

    def tower_forward(self, tower_inputs, other_outs=None):
        ys = []
        for h, tower, head in zip(tower_inputs, self.towers, self.output_layers):
            logit = tower(h)
            for extra in (other_outs or ()):
                logit = logit + extra
            ys.append(head(logit))
        return torch.cat(ys, dim=1)

    # -- regulariser
    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        weights = [weight_list] if isinstance(weight_list, nn.Parameter) else list(weight_list)
        self.regularization_weight.append((weights, l1, l2))

    def get_regularization_loss(self, device):
        from . import reg_ops
        if getattr(self, "_reg_folded", False):      # gradient applied inside FusedAdam: value only
            with torch.no_grad():
                return reg_ops.regularization_loss(self.regularization_weight, device)
        return reg_ops.regularization_loss(self.regularization_weight, device)

    def fold_regularization_into(self, optimizer):
        """Move the GRADIENT of the L2 regulariser into `optimizer` (optim.FusedAdam): the step adds
        2 * l2 * w itself (`g = scatter_part + (2 l2 + wd) W` on the fly), and
        get_regularization_loss() keeps returning the value, without an autograd path.  Same numbers as
        `loss + reg` followed by Adam; saves one table-sized pass and ~130 per-parameter gradient adds."""
        terms = {}
        for weights, l1, l2 in self.regularization_weight:
            if l1 > 0:
                raise ValueError("only L2 terms can be folded into the optimizer")
            for w in weights:
                p = w[1] if isinstance(w, tuple) else w
                if l2 > 0:
                    terms[p] = terms.get(p, 0.0) + float(l2)
        optimizer.fold_l2(terms)
        self._reg_folded = True

    def shard_table(self, group=None):
        """Row-shard the embedding table over the ranks of `group` (sharding.py); call after .to(device)
        and before building the optimizer.  The regulariser follows the new (shard) parameter."""
        old = self.embedding.embedding_dict.weight
        shards = self.embedding.shard_table(group)
        new = self.embedding.embedding_dict.weight
        self.regularization_weight = [([new if w is old else w for w in weights], l1, l2)
                                      for weights, l1, l2 in self.regularization_weight]
        return shards


def _weights_without_bn(module):
    """The reference's filter: named parameters whose name contains 'weight' but not 'bn'."""
    return [(n, p) for n, p in module.named_parameters() if "weight" in n and "bn" not in n]


# ---------------------------------------------------------------------------------------------
# Baseline-only layers (FactorizationMachine, DNN, CrossNetV2, CrossNetMix, ...) are out of scope
# (SURVEY.md 2, rows 17-19).  When the reference tree is reachable they are served from it so that
# the baseline models keep importing `model.layer`; nothing is copied into this repository.
# ---------------------------------------------------------------------------------------------
_REF_MODULE = None


def _reference_layer_module():
    global _REF_MODULE
    if _REF_MODULE is None:
        for root in (os.environ.get("AREAD_REF"), "/root/reference"):
            path = os.path.join(root, "model", "layer.py") if root else None
            if path and os.path.exists(path):
                spec = importlib.util.spec_from_file_location("_aread_reference_layer", path)
                module = importlib.util.module_from_spec(spec)
                sys.modules["_aread_reference_layer"] = module
                spec.loader.exec_module(module)
                _REF_MODULE = module
                break
    return _REF_MODULE


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    ref = _reference_layer_module()
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"model.layer has no attribute {name!r} (baseline-only layers are served from the "
                         "reference tree; set AREAD_REF to its location)")
