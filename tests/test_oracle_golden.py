"""Pins the oracle (oracle/) against the fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import re

import numpy as np
import pytest
import torch

from oracle import aread_torch as O
from oracle import embedding_np as E
from oracle import synth
from tests._util import CASES, assert_after_adam, assert_close, assert_compact, load_golden

# the oracle and the reference run the same torch CPU ops; only the association order of a few
# sums differs (e.g. mixing written as stack+sum), so the tolerance is a few fp32 ulps
RTOL, ATOL = 2e-5, 2e-6


# A Linear bias that feeds BatchNorm has an exactly-zero true gradient (BN removes the column
# mean): what Adam sees there is pure round-off noise, so only the hard bound applies.
PRE_BN_BIAS = re.compile(r"\.layers\.(0|4|8)\.bias$")


def _setup(name):
    fx = load_golden(name)
    spec = O.Spec(**fx["spec"])
    x, y = synth.random_batch(spec, fx["B"], seed=11, domain=fx["domain"], pad_id=fx["pad_id"])
    x2, y2 = synth.random_batch(spec, fx["B"], seed=12, domain=fx["domain"], pad_id=fx["pad_id"])
    return fx, spec, (x, y), (x2, y2)


@pytest.mark.parametrize("name", CASES)
def test_gather_bit_exact(name):
    fx, spec, (x, _), _ = _setup(name)
    sd = synth.deterministic_state(spec)
    got = E.gather_fwd(sd["embedding.embedding_dict.weight"].numpy(), x.numpy(), spec.offsets,
                       spec.flag, spec.seq_maxlen, spec.method)
    ref = fx["eval"]["embed"].numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # torch restatement used inside the float oracle agrees too
    assert torch.equal(O.embed(sd, spec, x), fx["eval"]["embed"])


@pytest.mark.parametrize("name", CASES)
def test_scatter_matches_reference_table_grad(name):
    """dense table gradient of sum(embed * G): reference order == np.add.at order, bit for bit"""
    fx, spec, (x, _), _ = _setup(name)
    sd = synth.deterministic_state(spec)
    W = sd["embedding.embedding_dict.weight"].clone().requires_grad_(True)
    out = O.embed({"embedding.embedding_dict.weight": W}, spec, x)
    G = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * G).sum().backward()
    got = E.scatter_bwd_dense(G.numpy(), x.numpy(), spec.offsets, spec.n_rows, spec.flag, spec.seq_maxlen, spec.method)
    assert np.array_equal(got.view(np.uint32), W.grad.numpy().view(np.uint32))
    # tiled order == sequential order when every segment sits inside one tile
    idx = E.lookup_rows(x.numpy(), spec.offsets, spec.n_rows).reshape(-1)
    g_cols = E.expand_pooled_grad(G.numpy(), x.shape[1], spec.flag, spec.seq_maxlen, spec.method)
    big = E.scatter_bwd_tiled(g_cols, idx, spec.n_rows, tile=1 << 20)
    assert np.array_equal(big.view(np.uint32), got.view(np.uint32))
    small = E.scatter_bwd_tiled(g_cols, idx, spec.n_rows, tile=8)
    np.testing.assert_allclose(small, got, rtol=1e-5, atol=1e-6)


def test_lookup_bounds_and_aliasing():
    off = E.field_offsets([5, 3, 4])
    assert off.tolist() == [0, 5, 8]
    off_mh = E.field_offsets([5, 3, 4], [False, False, False, True, True], itemid_idx=1)
    assert off_mh.tolist() == [0, 5, 8, 5, 5]
    x = np.array([[5, 0, 0]], dtype=np.int32)          # id 5 of field 0 aliases row 5 = field 1 row 0
    assert E.lookup_rows(x, off, 12).tolist() == [[5, 5, 8]]
    with pytest.raises(IndexError):
        E.lookup_rows(np.array([[0, 0, 4]], dtype=np.int32), off, 12)
    with pytest.raises(IndexError):
        E.lookup_rows(np.array([[-1, 0, 0]], dtype=np.int32), off, 12)
    assert E.gather_fwd(np.zeros((12, 4), np.float32), np.zeros((0, 3), np.int32), off).shape == (0, 3, 4)


def test_sort_segments_bookkeeping():
    rows = np.array([7, 2, 7, 7, 0, 2, 9])
    s, perm, uniq, seg = E.sort_segments(rows)
    assert s.tolist() == [0, 2, 2, 7, 7, 7, 9]
    assert perm.tolist() == [4, 1, 5, 0, 2, 3, 6]        # stable
    assert uniq.tolist() == [0, 2, 7, 9]
    assert seg.tolist() == [0, 1, 3, 6, 7]
    s, perm, uniq, seg = E.sort_segments(np.array([], dtype=np.int64))
    assert len(s) == 0 and seg.tolist() == [0]


@pytest.mark.parametrize("name", CASES)
def test_eval_forward_modes(name):
    fx, spec, (x, _), _ = _setup(name)
    sd = synth.deterministic_state(spec)
    ev = fx["eval"]
    with torch.no_grad():
        assert_close(O.forward(sd, spec, x, "wo_mask")["y"], ev["wo_mask"], RTOL, ATOL, "wo_mask")
        for mk, m in fx["masks"].items():
            out = O.forward(sd, spec, x, "domain_with_mask", m)
            assert_close(out["y"], ev[f"with_mask/{mk}"], RTOL, ATOL, f"with_mask/{mk}")
            out = O.forward(sd, spec, x, "domain_mask_bagging", m)
            assert_close(out["y"], ev[f"bagging/{mk}"], RTOL, ATOL, f"bagging/{mk}")
            for key, ref in ev[f"gate_means/{mk}"].items():
                assert_close(out["gate_means"][key], ref, RTOL, ATOL, f"gate_means/{mk}/{key}")
        out = O.forward(sd, spec, x[:1], "domain_with_mask", fx["masks"]["sparse"])
        assert_close(out["y"], ev["with_mask/b1"], RTOL, ATOL, "batch of one")
        assert_close(O.reg_loss(sd, spec), ev["reg"], 1e-6, 0, "reg")


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mk", ["full", "sparse"])
def test_train_steps(name, mk):
    fx, spec, b0, b1 = _setup(name)
    tr = fx[f"train/{mk}"]
    sd = O.make_leaf_params(synth.deterministic_state(spec))
    opt = O.make_adam(sd)
    mask = fx["masks"][mk]
    for step in range(3):
        xb, yb = b0 if step % 2 == 0 else b1
        if step == 0:
            out = O.forward(sd, spec, xb, "domain_mask_bagging", mask, training=True)
            data_loss = O.bagging_loss(out["y"], yb)
            reg = O.reg_loss(sd, spec)
            loss = data_loss + reg
            opt.zero_grad(set_to_none=True)
            loss.backward()
            assert_close(out["y"], tr["y_stack"], RTOL, ATOL, "y_stack")
            assert_close(data_loss, tr["data_loss"], RTOL, ATOL, "data loss")
            assert_close(reg, tr["reg"], 1e-6, 0, "reg")
            for key, ref in tr["gate_means"].items():
                assert_close(out["gate_means"][key], ref, RTOL, ATOL, f"gate_means/{key}")
            none_keys = sorted(k for k, v in sd.items() if v.requires_grad and v.grad is None)
            dead = ("atten_", "self_attns", "V_res")
            assert none_keys == [k for k in tr["grad_none"] if not k.startswith(dead)]
            for k, comp in tr["grads"].items():
                assert_compact(sd[k].grad, comp, 2e-4, 2e-7, f"grad {k}")
            opt.step()
        else:
            loss, _ = O.train_step(sd, spec, xb, yb, mask, opt)
        assert_close(loss, tr[f"loss{step}"], RTOL, ATOL, f"loss{step}")
    for k, comp in tr["state_after"].items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            assert_compact(sd[k], comp, 1e-4, 0.3 * 3 * 1e-3, f"state {k}")
        else:
            assert_after_adam(sd[k], comp, 3, 1e-3, f"state {k}", frac=1.0 if PRE_BN_BIAS.search(k) else 0.02)
    with torch.no_grad():
        out = O.forward(sd, spec, b0[0], "domain_with_mask", mask)
    assert_close(out["y"], tr["eval_after"], 1e-4, 1e-5, "eval after 3 steps")


@pytest.mark.parametrize("name", CASES)
def test_wo_mask_train(name):
    fx, spec, (x, y), _ = _setup(name)
    tr = fx["train/wo_mask"]
    sd = O.make_leaf_params(synth.deterministic_state(spec))
    out = O.forward(sd, spec, x, "wo_mask", training=True)
    loss = torch.nn.functional.binary_cross_entropy(out["y"].squeeze(), y.squeeze().float()) + O.reg_loss(sd, spec)
    loss.backward()
    assert_close(out["y"], tr["y"], RTOL, ATOL, "y")
    assert_close(loss, tr["loss"], RTOL, ATOL, "loss")
    for key, ref in tr["recorded"].items():
        assert_close(out["gates"][key].mean(dim=0), ref, RTOL, ATOL, f"recorded {key}")
    none_keys = sorted(k for k, v in sd.items() if v.requires_grad and v.grad is None)
    dead = ("atten_", "self_attns", "V_res")
    assert none_keys == [k for k in tr["grad_none"] if not k.startswith(dead)]
    for k, comp in tr["grads"].items():
        assert_compact(sd[k].grad, comp, 2e-4, 2e-7, f"grad {k}")
