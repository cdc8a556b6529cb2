"""The activation arena of the fused step (_mem.py): same numbers as allocator-backed buffers across changing
masks, ownership rules (two pending forwards, dropped graphs, retain_graph), and no growth once warm."""
import importlib

import pytest
import torch

from oracle import aread_torch as O
from oracle import synth
from tests._models import build_model
from tests._util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_mem = importlib.import_module("aread-multi-domain-recommendation_b200._mem")


def _setup(dropout=0.0):
    fx = load_golden("amz_small")
    spec = O.Spec(**fx["spec"])
    model = build_model(spec, DEV, dropout=dropout).train()
    model.expert_precision = "bf16x3"
    torch.manual_seed(0)
    masks = [synth.full_mask(spec)]
    for s in range(3):
        g = torch.Generator().manual_seed(s)
        m = [(torch.rand(t.shape, generator=g) < 0.6) for t in synth.full_mask(spec)]
        for t in m:                                   # every tower keeps at least one input edge
            t[0, :] = True
        masks.append(m)
    batches = [synth.random_batch(spec, fx["B"], seed=20 + i, domain=fx["domain"], pad_id=fx["pad_id"])
               for i in range(4)]
    return fx, model, masks, batches


def _step(model, fx, mask, batch):
    x = batch[0].to(DEV)
    y = batch[1].to(DEV).float().view(-1)
    preds = model(x, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=[m.to(DEV) for m in mask])
    loss = sum(torch.nn.functional.binary_cross_entropy(p, y) for p in preds.unbind(0)) / preds.shape[0]
    model.zero_grad()
    loss.backward()
    return preds.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


def test_arena_matches_allocator_across_masks(monkeypatch):
    fx, model, masks, batches = _setup()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    want = []
    monkeypatch.setattr(_mem, "ENABLED", False)
    for i in range(8):
        want.append(_step(model, fx, masks[i % 4], batches[i % 4]))
    model.load_state_dict(state)
    monkeypatch.setattr(_mem, "ENABLED", True)
    arena = model.arena(torch.device(DEV))
    for i in range(8):
        p, g = _step(model, fx, masks[i % 4], batches[i % 4])
        assert torch.equal(p, want[i][0]), f"step {i}"
        assert g.keys() == want[i][1].keys()
        for n in g:
            assert torch.equal(g[n], want[i][1][n]), f"step {i} grad {n}"
        assert not arena.busy
    assert arena.cap > 0 and arena.need <= arena.cap          # warm: everything came out of the arena
    cap = arena.cap
    for i in range(4):
        _step(model, fx, masks[i], batches[i])
    assert arena.cap == cap


def test_two_pending_forwards_and_dropped_graph():
    fx, model, masks, batches = _setup()
    x0, x1 = batches[0][0].to(DEV), batches[1][0].to(DEV)
    full = [m.to(DEV) for m in masks[0]]
    for _ in range(2):                                         # warm the arena
        _step(model, fx, masks[0], batches[0])
    arena = model.arena(torch.device(DEV))
    want0 = _step(model, fx, masks[0], batches[0])[1]
    want1 = _step(model, fx, masks[0], batches[1])[1]
    p0 = model(x0, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=full)        # owns the arena
    assert arena.busy
    p1 = model(x1, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=full)        # allocator-backed
    y0, y1 = batches[0][1].to(DEV).float().view(-1), batches[1][1].to(DEV).float().view(-1)
    bce = torch.nn.functional.binary_cross_entropy
    model.zero_grad()
    (sum(bce(p, y1) for p in p1.unbind(0)) / p1.shape[0]).backward()
    got1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    (sum(bce(p, y0) for p in p0.unbind(0)) / p0.shape[0]).backward()
    got0 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for n in want0:
        assert torch.equal(got0[n], want0[n]) and torch.equal(got1[n], want1[n]), n
    assert not arena.busy
    p2 = model(x0, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=full)        # graph dropped without a backward
    assert arena.busy
    del p2
    assert not arena.busy
    with torch.no_grad():
        model(x0, mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=full)
    assert not arena.busy


def test_second_backward_is_refused():
    fx, model, masks, batches = _setup()
    for _ in range(2):
        _step(model, fx, masks[0], batches[0])
    full = [m.to(DEV) for m in masks[0]]
    p = model(batches[0][0].to(DEV), mode="domain_mask_bagging", domain_i=fx["domain"], current_mask=full)
    loss = p.mean()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        loss.backward()


def test_folded_regulariser_trains_like_the_loss_term():
    """BaseModel.fold_regularization_into: value still in the loss, gradient applied by FusedAdam."""
    optim = importlib.import_module("aread-multi-domain-recommendation_b200.optim")
    fx, model_a, masks, batches = _setup()
    _, model_b, _, _ = _setup()
    kw = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    opt_a, opt_b = optim.FusedAdam(model_a.parameters(), **kw), optim.FusedAdam(model_b.parameters(), **kw)
    model_b.fold_regularization_into(opt_b)
    for i in range(6):
        losses = []
        for model, opt in ((model_a, opt_a), (model_b, opt_b)):
            x, y = batches[i % 4][0].to(DEV), batches[i % 4][1].to(DEV)
            preds = model(x, mode="domain_mask_bagging", domain_i=fx["domain"],
                          current_mask=[m.to(DEV) for m in masks[i % 4]])
            loss = model.bagging_loss(preds, y) + model.get_regularization_loss(device=torch.device(DEV))
            model.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.detach())
        torch.testing.assert_close(losses[1], losses[0], rtol=1e-5, atol=1e-6)
    sa, sb = model_a.state_dict(), model_b.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point:
            torch.testing.assert_close(sb[k], sa[k], rtol=2e-5, atol=2e-6, msg=k)
