"""CUDA-graph replay of the fused node (fused.py): recorded forward / backward sequences give bit-identical results
to the eager launches -- across alternating masks, with dropout (seed read from device memory), in eval mode, under
gradient accumulation (falls back to eager launches) and after the weights were reloaded in place."""
import importlib

import pytest
import torch

from tests.test_arena_gpu import DEV, _setup, _step

pytestmark = pytest.mark.gpu
fused = importlib.import_module("aread-multi-domain-recommendation_b200.fused")
optim = importlib.import_module("aread-multi-domain-recommendation_b200.optim")


def _on_device(masks):
    """Masks as the model keeps them (device tensors that persist from step to step, like domain_mask[d])."""
    return [[m.to(DEV) for m in mk] for mk in masks]


@pytest.fixture(autouse=True)
def _record_candidates_early(monkeypatch):
    """These tests hand their masks in as `current_mask=`; record them as early as installed domain masks."""
    monkeypatch.setattr(fused, "GRAPH_AFTER_CANDIDATE", fused.GRAPH_AFTER)


def _run(model, fx, masks, batches, n, graphs, monkeypatch, train=True):
    monkeypatch.setattr(fused, "USE_GRAPHS", graphs)
    out = []
    for i in range(n):
        torch.manual_seed(1000 + i)                        # the dropout seed of the step comes from the CPU generator
        out.append(_step(model, fx, masks[i % 2], batches[i % 2]))
    return out


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_replay_matches_eager(monkeypatch, dropout):
    fx, model, masks, batches = _setup(dropout=dropout)
    masks = _on_device(masks)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    want = _run(model, fx, masks, batches, 10, False, monkeypatch)
    model.load_state_dict(state)
    got = _run(model, fx, masks, batches, 10, True, monkeypatch)
    recorded = [e for e in model._graphs.entries.values() if e.fwd is not None]
    assert len(recorded) == 2 and all(e.bwd is not None for e in recorded)
    for i, ((p0, g0), (p1, g1)) in enumerate(zip(want, got)):
        assert torch.equal(p1, p0), f"step {i}"
        assert g1.keys() == g0.keys()
        for n in g0:
            assert torch.equal(g1[n], g0[n]), f"step {i} grad {n}"
    sd = model.state_dict()                                 # running statistics advanced identically
    model.load_state_dict(state)
    _run(model, fx, masks, batches, 10, False, monkeypatch)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_training_with_replay_matches_eager(monkeypatch):
    res = []
    for graphs in (False, True):
        fx, model, masks, batches = _setup(dropout=0.2)
        masks = _on_device(masks)
        monkeypatch.setattr(fused, "USE_GRAPHS", graphs)
        opt = optim.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        model.fold_regularization_into(opt)
        for i in range(8):
            torch.manual_seed(50 + i)
            x, y = batches[i % 2][0].to(DEV), batches[i % 2][1].to(DEV)
            preds = model(x, mode="domain_mask_bagging", domain_i=fx["domain"],
                          current_mask=masks[i % 2])
            loss = model.bagging_loss(preds, y) + model.get_regularization_loss(device=torch.device(DEV))
            model.zero_grad()
            loss.backward()
            opt.step()
        res.append({k: v.clone() for k, v in model.state_dict().items()})
        if graphs:
            assert any(e.bwd is not None for e in model._graphs.entries.values())
    for k in res[0]:
        assert torch.equal(res[1][k], res[0][k]), k


def test_accumulation_eval_and_reload(monkeypatch):
    fx, model, masks, batches = _setup()
    masks = _on_device(masks)
    monkeypatch.setattr(fused, "USE_GRAPHS", True)
    for i in range(5):                                      # records the sequences of mask 0
        _step(model, fx, masks[0], batches[0])
    entry = next(e for e in model._graphs.entries.values() if e.fwd is not None)
    assert entry.bwd is not None
    p, g = _step(model, fx, masks[0], batches[0])
    # second backward into existing .grad: eager launches, gradients add up
    x, y = batches[0][0].to(DEV), batches[0][1].to(DEV).float().view(-1)
    preds = model(x, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=masks[0])
    loss = sum(torch.nn.functional.binary_cross_entropy(q, y) for q in preds.unbind(0)) / preds.shape[0]
    loss.backward()
    for n, prm in model.named_parameters():
        if prm.grad is not None and n in g:
            torch.testing.assert_close(prm.grad, 2 * g[n], rtol=1e-6, atol=1e-9, msg=n)
    # eval mode has its own sequence; compare with eager
    model.eval()
    mk = masks[0]
    with torch.no_grad():
        outs = [model(x, mode="domain_with_mask", domain_i=fx["domain"], current_mask=mk).clone() for _ in range(5)]
        monkeypatch.setattr(fused, "USE_GRAPHS", False)
        ref = model(x, mode="domain_with_mask", domain_i=fx["domain"], current_mask=mk)
        monkeypatch.setattr(fused, "USE_GRAPHS", True)
    assert all(torch.equal(o, ref) for o in outs)
    # weights reloaded in place: the recorded sequences read the new values
    sd = {k: (v * 0.5 if v.dtype.is_floating_point else v) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    with torch.no_grad():
        new = model(x, mode="domain_with_mask", domain_i=fx["domain"], current_mask=mk)
        monkeypatch.setattr(fused, "USE_GRAPHS", False)
        ref = model(x, mode="domain_with_mask", domain_i=fx["domain"], current_mask=mk)
    assert torch.equal(new, ref) and not torch.equal(new, outs[0])


def test_record_graphs_ahead_of_time(monkeypatch):
    fx, model, masks, batches = _setup(dropout=0.2)
    monkeypatch.setattr(fused, "USE_GRAPHS", True)
    dev_masks = _on_device(masks)
    domains = [0, 1, 2]
    for d in domains:
        model.domain_mask[d] = dev_masks[d]
    x, y = batches[0][0].to(DEV), batches[0][1].to(DEV)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(9)
    rng = torch.get_rng_state()
    model.record_graphs(x, domains=domains)
    assert torch.equal(torch.get_rng_state(), rng)
    for k, v in model.state_dict().items():
        assert torch.equal(v, state[k]), k
    assert all(p.grad is None for p in model.parameters())
    recorded = [e for e in model._graphs.entries.values() if e.fwd is not None and e.bwd is not None]
    assert len(recorded) == len(domains)
    calls = [e.calls for e in recorded]
    # the very next train step of each domain replays, and equals the eager launches
    got = []
    for d in domains:
        torch.manual_seed(70 + d)
        preds = model(x, mode="domain_mask_bagging", domain_i=d)
        loss = model.bagging_loss(preds, y)
        model.zero_grad()
        loss.backward()
        got.append((preds.detach().clone(), model.embedding.embedding_dict.weight.grad.clone()))
    assert [e.calls for e in recorded] == [c + 1 for c in calls]
    model.load_state_dict(state)
    monkeypatch.setattr(fused, "USE_GRAPHS", False)
    for d, (p, g) in zip(domains, got):
        torch.manual_seed(70 + d)
        preds = model(x, mode="domain_mask_bagging", domain_i=d)
        loss = model.bagging_loss(preds, y)
        model.zero_grad()
        loss.backward()
        assert torch.equal(preds.detach(), p) and torch.equal(model.embedding.embedding_dict.weight.grad, g)


def _masks_of_different_size(model):
    """HEMP masks whose numbers of active towers differ (workspaces and compact activations change size)."""
    import numpy as np
    out = []
    for seed, p in ((3, 0.15), (4, 0.95), (5, 0.4), (6, 0.25)):
        np.random.seed(seed)
        out.append(model.generate_mask("rand", 0, init_active_percent=p))
    sizes = {tuple(len(a) for a in model.mask_info(m).active_idx) for m in out}
    assert len(sizes) >= 3, sizes
    return out


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_interleaved_masks_of_different_size(monkeypatch, dropout):
    """A sequence recorded under a small mask keeps giving the eager results after larger masks ran in between
    (their larger workspaces / arena growth must not leave the recorded sequence pointing at freed memory)."""
    fx, model, _, batches = _setup(dropout=dropout)
    masks = _masks_of_different_size(model)
    order = [0, 0, 0, 0, 1, 0, 1, 1, 1, 0, 2, 0, 3, 3, 3, 3, 1, 0, 2, 3]
    state = {k: v.clone() for k, v in model.state_dict().items()}

    def run(graphs):
        monkeypatch.setattr(fused, "USE_GRAPHS", graphs)
        model.load_state_dict(state)
        res = []
        for i, mi in enumerate(order):
            torch.manual_seed(300 + i)
            x = batches[i % 4][0].to(DEV)
            y = batches[i % 4][1].to(DEV).float().view(-1)
            preds = model(x, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=masks[mi])
            loss = sum(torch.nn.functional.binary_cross_entropy(p, y) for p in preds.unbind(0)) / preds.shape[0]
            model.zero_grad()
            loss.backward()
            res.append((preds.detach().clone(),
                        {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
        return res

    want = run(False)
    got = run(True)
    assert sum(e.fwd is not None for e in model._graphs.entries.values()) >= 3
    for i, ((p0, g0), (p1, g1)) in enumerate(zip(want, got)):
        assert torch.equal(p1, p0), f"step {i} (mask {order[i]})"
        assert g1.keys() == g0.keys()
        for n in g0:
            assert torch.equal(g1[n], g0[n]), f"step {i} grad {n}"


def test_two_pending_forwards_with_dropout_after_recording(monkeypatch):
    """Once a sequence is recorded, a second forward issued before the first one's backward (it cannot lease the
    arena and runs eagerly) must not disturb the dropout seed the first backward regenerates its masks from."""
    fx, model, masks, batches = _setup(dropout=0.2)
    mask = _on_device(masks)[0]
    model.domain_mask[fx["domain"]] = mask
    monkeypatch.setattr(fused, "USE_GRAPHS", True)
    x0, x1 = batches[0][0].to(DEV), batches[1][0].to(DEV)
    y0 = batches[0][1].to(DEV)

    def grads_of(two_pending):
        torch.manual_seed(77)
        p0 = model(x0, mode="domain_mask_bagging", domain_i=fx["domain"])
        if two_pending:
            p1 = model(x1, mode="domain_mask_bagging", domain_i=fx["domain"])      # eager, own seed by value
            with torch.no_grad():
                model(x1, mode="domain_mask_bagging", domain_i=fx["domain"])       # and a no_grad train-mode forward
        model.zero_grad()
        model.bagging_loss(p0, y0).backward()
        g = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        if two_pending:
            model.zero_grad()
            p1.sum().backward()
        return p0.detach().clone(), g

    for _ in range(4):                                                          # records forward and backward
        grads_of(False)
    assert any(e.bwd is not None for e in model._graphs.entries.values())
    buffers = {n: b.clone() for n, b in model.named_buffers()}
    p_ref, g_ref = grads_of(False)
    for n, b in model.named_buffers():
        b.copy_(buffers[n])
    p_got, g_got = grads_of(True)
    assert torch.equal(p_got, p_ref)
    for n in g_ref:
        assert torch.equal(g_got[n], g_ref[n]), n
