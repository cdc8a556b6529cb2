"""Deterministic model states and inputs for parity runs (TEST INFRASTRUCTURE, see
oracle/__init__.py).

Weights are a pure function of (state_dict key, shape): the golden generator (which imports
the reference in the build container) and the GPU-box tests (which cannot) rebuild the same
state without shipping megabytes of tensors.
"""
import zlib

import numpy as np
import torch

from .aread_torch import Spec


def state_shapes(spec: Spec, with_attention=False, atten_embed_dim=64, att_layer_num=3):
    """{key: shape} of the reference state_dict for `spec`.  Checked against the
    real reference module in tests/golden/make_golden.py."""
    D, E = spec.embed_dim, spec.E
    s = {}
    s["embedding.embedding_dict.weight"] = (spec.n_rows, D)
    s["linear.fc.weight"] = (1, E)
    s["linear.fc.bias"] = (1,)
    s["group_embedding.weight"] = (spec.n_tower[0], D)
    s["final_gate.0.weight"] = (spec.n_tower[-1], 2 * D)
    for k in range(spec.n_cross_layers):
        s[f"cn.w.{k}.weight"] = (1, E)
    for k in range(spec.n_cross_layers):
        s[f"cn.b.{k}"] = (E,)
    if with_attention:
        A = atten_embed_dim
        s["atten_embedding.weight"] = (A, D)
        s["atten_embedding.bias"] = (A,)
        for i in range(att_layer_num):
            s[f"self_attns.{i}.in_proj_weight"] = (3 * A, A)
            s[f"self_attns.{i}.in_proj_bias"] = (3 * A,)
            s[f"self_attns.{i}.out_proj.weight"] = (A, A)
            s[f"self_attns.{i}.out_proj.bias"] = (A,)
        s["V_res_embedding.weight"] = (A, D)
        s["V_res_embedding.bias"] = (A,)
        s["atten_linear.weight"] = (1, spec.out_fields * A)

    def add_mlp(prefix, d_in, dims):
        for i, d_out in enumerate(dims):
            li = 4 * i
            s[f"{prefix}.layers.{li}.weight"] = (d_out, d_in)
            s[f"{prefix}.layers.{li}.bias"] = (d_out,)
            bn = f"{prefix}.layers.{li + 1}"
            s[bn + ".weight"] = (d_out,)
            s[bn + ".bias"] = (d_out,)
            s[bn + ".running_mean"] = (d_out,)
            s[bn + ".running_var"] = (d_out,)
            s[bn + ".num_batches_tracked"] = ()
            d_in = d_out

    for k in range(spec.n_expert):
        add_mlp(f"mmoe_experts.{k}", E, spec.expert_dims)
    for g in range(spec.n_tower[0]):
        s[f"mmoe_gates.{g}.0.weight"] = (spec.n_expert, E)
        s[f"mmoe_gates.{g}.0.bias"] = (spec.n_expert,)
    d_in = spec.expert_dims[-1]
    for l in range(spec.n_level):
        for t in range(spec.n_tower[l]):
            add_mlp(f"towers.{l}.{t}", d_in, spec.tower_dims[l])
        d_in = spec.tower_dims[l][-1]
    for l in range(1, spec.n_level):
        for t in range(spec.n_tower[l]):
            s[f"tower_gates.{l - 1}.{t}.0.weight"] = (spec.n_tower[l - 1], 2 * D)
            s[f"tower_gates.{l - 1}.{t}.0.bias"] = (spec.n_tower[l - 1],)
    for t in range(spec.n_tower[-1]):
        s[f"towers_linear.{t}.weight"] = (1, E + spec.tower_dims[-1][-1])
    return s


def _gen(key):
    g = torch.Generator()
    g.manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    return g


def deterministic_tensor(key, shape):
    """Value of state entry `key`: N(0,1) tables, fan-in scaled linears, BN affine near identity,
    non-trivial running statistics so that eval-mode BN is exercised."""
    g = _gen(key)
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.tensor(3, dtype=torch.long)
    if key.endswith("running_mean"):
        return 0.1 * torch.randn(shape, generator=g)
    if key.endswith("running_var"):
        return 0.5 + torch.rand(shape, generator=g)
    if key in ("embedding.embedding_dict.weight", "group_embedding.weight"):
        return torch.randn(shape, generator=g)
    if key.startswith("cn.b."):
        return 0.05 * torch.randn(shape, generator=g)
    is_bn = len(shape) == 1 and (".layers." in key) and (int(key.split(".layers.")[1].split(".")[0]) % 4 == 1)
    if is_bn and key.endswith(".weight"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if key.endswith(".bias"):
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) == 2:
        return torch.randn(shape, generator=g) / float(np.sqrt(shape[1]))
    return 0.1 * torch.randn(shape, generator=g)


def deterministic_state(spec: Spec, with_attention=False):
    return {k: deterministic_tensor(k, shp) for k, shp in state_shapes(spec, with_attention).items()}


def random_batch(spec: Spec, B, seed, domain=None, pad_id=None, pad_prob=0.4):
    """Seeded [B, n_cols] int32 ids + [B, 1] int16 labels.  One-hot ids are uniform in their
    field; multi-hot columns draw item ids (the item field's range) with trailing `pad_id`s."""
    rng = np.random.RandomState(seed)
    dims = np.asarray(spec.one_hot_field_dims, dtype=np.int64)
    cols = [rng.randint(0, d, size=B) for d in dims]
    if domain is not None:
        cols[spec.domain_idx] = np.full(B, domain)
    for f in range(spec.n_mh_fields):
        seq = rng.randint(0, dims[spec.itemid_idx], size=(B, spec.seq_maxlen))
        if pad_id is not None:
            n_real = rng.binomial(spec.seq_maxlen, 1 - pad_prob, size=B)
            pos = np.arange(spec.seq_maxlen)[None, :]
            seq = np.where(pos < n_real[:, None], seq, pad_id)
        cols.extend(list(seq.T))
    x = torch.from_numpy(np.stack(cols, axis=1).astype(np.int32))
    y = torch.from_numpy(rng.binomial(1, 0.3, size=(B, 1)).astype(np.int16))
    return x, y


def full_mask(spec: Spec, value=True):
    nt = spec.n_tower
    shapes = [(1, nt[0])] + [(nt[l - 1], nt[l]) for l in range(1, len(nt))] + [(nt[-1], 1)]
    return [torch.full(s, bool(value), dtype=torch.bool) for s in shapes]
