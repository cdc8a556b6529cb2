"""Parity of the CUDA embedding path (through the C ABI) with the oracle.  Bit-exact: the gather,
the pooled sums, the sort/segment bookkeeping and the documented scatter summation order."""
import importlib

import numpy as np
import pytest
import torch

from oracle import aread_torch as O
from oracle import embedding_np as E
from oracle import synth
from tests._util import CASES, load_golden

pytestmark = pytest.mark.gpu

PKG = importlib.import_module("aread-multi-domain-recommendation_b200")
ops = importlib.import_module("aread-multi-domain-recommendation_b200.embedding_ops")
layer = importlib.import_module("aread-multi-domain-recommendation_b200.layer")
DEV = "cuda:0"


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_embedding(dims, D, flag=None, itemid_idx=0, L=1, method=None, seed=0):
    mh = {"multi_hot_flag": list(flag) if flag is not None else [False] * len(dims), "itemid_idx": itemid_idx,
          "seq_maxlen": L, "method": method}
    emb = layer.FeaturesEmbedding(np.asarray(dims), D, mh)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        emb.embedding_dict.weight.copy_(torch.randn(emb.embedding_dict.weight.shape, generator=g))
    return emb.to(DEV)


def random_ids(dims, n_mh_fields, L, itemid_idx, B, seed, pad_id=None):
    rng = np.random.RandomState(seed)
    cols = [rng.randint(0, d, size=B) for d in dims]
    for _ in range(n_mh_fields):
        seq = rng.randint(0, dims[itemid_idx], size=(B, L))
        if pad_id is not None:
            seq = np.where(rng.rand(B, L) < 0.4, pad_id, seq)
        cols.extend(list(seq.T))
    return np.stack(cols, axis=1).astype(np.int32) if B else np.zeros((0, len(dims) + n_mh_fields * L), np.int32)


@pytest.mark.parametrize("name", CASES)
def test_gather_matches_golden(name):
    fx = load_golden(name)
    spec = O.Spec(**fx["spec"])
    x, _ = synth.random_batch(spec, fx["B"], seed=11, domain=fx["domain"], pad_id=fx["pad_id"])
    emb = make_embedding(spec.one_hot_field_dims, spec.embed_dim, spec.flag, spec.itemid_idx, spec.seq_maxlen,
                         spec.method)
    with torch.no_grad():
        emb.embedding_dict.weight.copy_(synth.deterministic_tensor("embedding.embedding_dict.weight",
                                                                   (spec.n_rows, spec.embed_dim)))
        out = emb(x.to(DEV))
    torch.cuda.synchronize()
    assert np.array_equal(bits(out.cpu().numpy()), bits(fx["eval"]["embed"].numpy()))


@pytest.mark.parametrize("D", [4, 8, 12, 16, 32, 64, 128])
@pytest.mark.parametrize("method,L", [(None, 1), ("mean", 5), ("sum", 3)])
@pytest.mark.parametrize("B", [0, 1, 7, 1000])
def test_gather_bit_exact_shapes(D, method, L, B):
    dims = [50, 7, 300, 11]
    n_mh = 2 if method else 0
    flag = [False] * 4 + [True] * (n_mh * L)
    emb = make_embedding(dims, D, flag, itemid_idx=2, L=L, method=method, seed=D)
    x = random_ids(dims, n_mh, L, 2, B, seed=B + D, pad_id=300 if method else None)
    with torch.no_grad():
        out = emb(torch.from_numpy(x).to(DEV))
    ref = E.gather_fwd(emb.embedding_dict.weight.detach().cpu().numpy(), x, emb.offsets, np.array(flag), L, method)
    assert tuple(out.shape) == ref.shape
    assert np.array_equal(bits(out.cpu().numpy()), bits(ref))


def test_gather_bf16_copy_is_rne():
    dims = [97, 13]
    emb = make_embedding(dims, 32)
    x = torch.from_numpy(random_ids(dims, 0, 1, 0, 513, seed=3)).to(DEV)
    plan = emb.plan(torch.device(DEV))
    out, out_bf16 = ops.gather(plan, emb.embedding_dict.weight.detach(), x, want_bf16=True)
    assert torch.equal(out_bf16, out.flatten(1).to(torch.bfloat16))


def test_out_of_range_raises_index_error(monkeypatch):
    monkeypatch.setattr(layer, "BOUNDS_MODE", "sync")
    dims = [10, 5]
    emb = make_embedding(dims, 32)
    ok = torch.tensor([[10, 0]], dtype=torch.int32, device=DEV)        # aliases into field 1: allowed
    with torch.no_grad():
        out = emb(ok)
    assert torch.equal(out[0, 0], emb.embedding_dict.weight[10])
    with pytest.raises(IndexError):
        with torch.no_grad():
            emb(torch.tensor([[0, 5]], dtype=torch.int32, device=DEV))   # row 15 of 15
    with pytest.raises(IndexError):
        with torch.no_grad():
            emb(torch.tensor([[-1, 0]], dtype=torch.int32, device=DEV))
    with torch.no_grad():                                               # flag was cleared, next call is fine
        emb(ok)


def test_cpu_tensors_are_rejected():
    emb = make_embedding([10, 5], 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        emb(torch.zeros((2, 2), dtype=torch.int32))


def _scatter_case(dims, D, flag, itemid_idx, L, method, B, seed, skew=False):
    emb = make_embedding(dims, D, flag, itemid_idx, L, method, seed=seed)
    n_mh = (int(np.sum(flag)) // L) if method else 0
    x = random_ids(dims, n_mh, L, itemid_idx, B, seed=seed + 1)
    if skew:
        x[:, 1] = 3                                                     # one row hit by every sample
    plan = emb.plan(torch.device(DEV))
    g = torch.randn((B, plan.n_fields, D), generator=torch.Generator().manual_seed(seed + 2))
    return emb, plan, x, g


@pytest.mark.parametrize("D", [8, 32, 64])
@pytest.mark.parametrize("method,L", [(None, 1), ("mean", 5)])
def test_scatter_documented_order_bit_exact(D, method, L):
    dims = [40, 6, 500, 9]
    flag = [False] * 4 + [True] * ((2 * L) if method else 0)
    B = 600
    emb, plan, x, g = _scatter_case(dims, D, flag, 2, L, method, B, seed=D, skew=True)
    dw, rows, pos = ops.scatter(plan, torch.from_numpy(x).to(DEV), g.to(DEV), want_sorted=True)
    torch.cuda.synchronize()
    idx = E.lookup_rows(x, emb.offsets, plan.n_rows).reshape(-1)
    s_rows, perm, _, _ = E.sort_segments(idx)
    assert np.array_equal(rows.cpu().numpy(), s_rows.astype(np.int32))       # sorted rows
    assert np.array_equal(pos.cpu().numpy(), perm.astype(np.int32))          # stable permutation
    g_cols = E.expand_pooled_grad(g.numpy(), x.shape[1], np.array(flag), L, method, reciprocal=True)
    ref = E.scatter_bwd_tiled(g_cols, idx, plan.n_rows)
    assert np.array_equal(bits(dw.cpu().numpy()), bits(ref))
    # and it is the reference gradient up to summation order
    seq = E.scatter_bwd_dense(g.numpy(), x, emb.offsets, plan.n_rows, np.array(flag), L, method)
    np.testing.assert_allclose(dw.cpu().numpy(), seq, rtol=2e-5, atol=2e-5)


def test_scatter_documented_order_four_levels():
    """1.3e5 lookups -> a 4-level tree; one row is hit by every sample, others by thousands"""
    dims = [3, 40, 70000]
    B = 44000
    emb, plan, x, g = _scatter_case(dims, 32, [False] * 3, 0, 1, None, B, seed=9, skew=True)
    dw = ops.scatter(plan, torch.from_numpy(x).to(DEV), g.to(DEV))
    idx = E.lookup_rows(x, emb.offsets, plan.n_rows).reshape(-1)
    ref = E.scatter_bwd_tiled(g.numpy().reshape(-1, 32), idx, plan.n_rows)
    assert np.array_equal(bits(dw.cpu().numpy()), bits(ref))


def test_scatter_equals_reference_order_without_long_segments():
    """all segments fit one tile -> the sequential reference order, bit for bit"""
    dims = [100000, 90000]
    B = 300
    emb = make_embedding(dims, 32)
    rng = np.random.RandomState(0)
    x = np.stack([rng.permutation(dims[0])[:B], rng.permutation(dims[1])[:B]], axis=1).astype(np.int32)
    x[5:9, 0] = x[4, 0]                                                     # a few short duplicate runs
    plan = emb.plan(torch.device(DEV))
    g = torch.randn((B, 2, 32), generator=torch.Generator().manual_seed(1))
    dw = ops.scatter(plan, torch.from_numpy(x).to(DEV), g.to(DEV))
    seq = E.scatter_bwd_dense(g.numpy(), x, emb.offsets, plan.n_rows)
    got = dw.cpu().numpy()
    touched = np.unique(E.lookup_rows(x, emb.offsets, plan.n_rows))
    # a segment may still straddle a tile boundary; those rows are compared to tolerance
    exact = sum(np.array_equal(bits(got[r]), bits(seq[r])) for r in touched)
    assert exact >= len(touched) - 2 * (x.size // 32 + 1)
    np.testing.assert_allclose(got, seq, rtol=1e-6, atol=1e-6)


def test_scatter_is_deterministic_and_autograd_wired():
    fx = load_golden("amz_small")
    spec = O.Spec(**fx["spec"])
    x, _ = synth.random_batch(spec, 4096, seed=5, domain=2, pad_id=500)
    emb = make_embedding(spec.one_hot_field_dims, spec.embed_dim, spec.flag, spec.itemid_idx, spec.seq_maxlen,
                         spec.method)
    xg = x.to(DEV)
    G = torch.randn((4096, spec.out_fields, spec.embed_dim), device=DEV)
    grads = []
    for _ in range(3):
        emb.zero_grad()
        (emb(xg) * G).sum().backward()
        grads.append(emb.embedding_dict.weight.grad.clone())
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[1], grads[2])
    W = emb.embedding_dict.weight.detach().cpu().clone().requires_grad_(True)
    (O.embed({"embedding.embedding_dict.weight": W}, spec, x) * G.cpu()).sum().backward()
    torch.testing.assert_close(grads[0].cpu(), W.grad, rtol=1e-4, atol=1e-3)   # row 500 sums ~10^4 terms


def test_full_size_properties():
    """AliCCP-scale vocabularies, batch 65536: gather against torch indexing (bit-exact), scatter
    through size-independent properties (column checksums, untouched rows stay zero, determinism)."""
    dims = [211161, 95, 14, 3, 8, 4, 4, 3, 5, 41775, 30, 284915, 81491, 112993, 1929, 118091, 54472, 34677, 5821,
            106908, 54295, 31716, 4]
    B = 65536
    emb = make_embedding(dims, 32)
    rng = np.random.RandomState(7)
    x = np.stack([np.minimum((rng.zipf(1.3, size=B) - 1) % d, d - 1) for d in dims], axis=1).astype(np.int32)
    x[:, 10] = 4
    xg = torch.from_numpy(x).to(DEV)
    plan = emb.plan(torch.device(DEV))
    W = emb.embedding_dict.weight.detach()
    out, _ = ops.gather(plan, W, xg)
    idx = xg.long() + torch.from_numpy(emb.offsets).to(DEV)
    assert torch.equal(out, W[idx])
    g = torch.randn((B, len(dims), 32), device=DEV)
    dw1 = ops.scatter(plan, xg, g).clone()
    dw2 = ops.scatter(plan, xg, g)
    assert torch.equal(dw1, dw2)
    touched = torch.zeros(plan.n_rows, dtype=torch.bool, device=DEV)
    touched[idx.reshape(-1)] = True
    assert not dw1[~touched].any()
    torch.testing.assert_close(dw1.double().sum(dim=0), g.double().sum(dim=(0, 1)), rtol=1e-6, atol=1e-3)
    ref = torch.zeros_like(dw1).index_add_(0, idx.reshape(-1), g.reshape(-1, 32))
    torch.testing.assert_close(dw1, ref, rtol=1e-4, atol=1e-2)
