"""SURVEY 8(f) rank 2 -- the HEMP regroup loop's state handling: `save_model_state` / `load_model_state`
(reference model/aread.py:534-546: deep copy + load_state_dict of every entry under the roll-back prefixes) as
persistent snapshot buffers + one multi-tensor copy launch, and `FusedAdam.reset()` standing in for the fresh
`optimizer_fast` of run.py:632-633.  Results must equal the reference semantics exactly (they are copies)."""
import importlib
import re

import pytest
import torch

from oracle import aread_torch as O
from tests._models import build_model
from tests._util import load_golden
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
optim = importlib.import_module("aread-multi-domain-recommendation_b200.optim")


def _model():
    fx = load_golden("ali_small")
    spec = O.Spec(**fx["spec"])
    return fx, spec, build_model(spec, DEV, dropout=0.0).train()


def _train(model, fx, spec, opt, steps, seed):
    mask = [m.to(DEV) for m in fx["masks"]["sparse"]]
    for i in range(steps):
        x, y = synth.random_batch(spec, 64, seed=seed + i, domain=fx["domain"])
        preds = model(x.to(DEV), mode="domain_mask_bagging", current_mask=mask)
        loss = model.bagging_loss(preds, y.to(DEV)) + model.get_regularization_loss(device=torch.device(DEV))
        model.zero_grad()
        loss.backward()
        opt.step()


def test_rollback_restores_exactly_the_reference_prefixes():
    fx, spec, model = _model()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    _train(model, fx, spec, opt, 2, 0)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model.save_model_state()
    snap_ptrs = {k: v.data_ptr() for k, v in model.model_state.items()}
    pattern = re.compile('^(' + '|'.join(model._ROLLBACK_PREFIXES) + ')')
    assert set(model.model_state) == {k for k in before if pattern.match(k)}
    assert not any(k.startswith(("mmoe_experts", "mmoe_gates", "group_embedding", "final_gate")) for k in model.model_state)
    for k, v in model.model_state.items():
        assert torch.equal(v, before[k]) and v.data_ptr() != model.state_dict()[k].data_ptr()
    _train(model, fx, spec, opt, 3, 10)
    moved = {k: v.clone() for k, v in model.state_dict().items()}
    assert not torch.equal(moved["embedding.embedding_dict.weight"], before["embedding.embedding_dict.weight"])
    model.load_model_state()
    for k, v in model.state_dict().items():
        want = before[k] if pattern.match(k) else moved[k]        # experts / MMoE gates are NOT rolled back
        assert torch.equal(v, want), k
    # a second save reuses the snapshot buffers (no table-sized allocation per regroup) and sees the new values
    _train(model, fx, spec, opt, 1, 20)
    model.save_model_state()
    assert {k: v.data_ptr() for k, v in model.model_state.items()} == snap_ptrs
    for k, v in model.model_state.items():
        assert torch.equal(v, model.state_dict()[k]), k
    # storage replaced behind the snapshot's back (model.to / a new table): the restore still lands in the live tensors
    model.embedding.embedding_dict.weight = torch.nn.Parameter(model.embedding.embedding_dict.weight.detach().clone() + 1)
    model.load_model_state()
    assert torch.equal(model.embedding.embedding_dict.weight, model.model_state["embedding.embedding_dict.weight"])


def test_fused_adam_reset_equals_fresh_optimizer():
    fx, spec, model = _model()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    kw = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    fresh_runs = []
    for _ in range(2):
        model.load_state_dict(state)
        _train(model, fx, spec, optim.FusedAdam(model.parameters(), **kw), 3, 40)
        fresh_runs.append({k: v.clone() for k, v in model.state_dict().items()})
    opt = optim.FusedAdam(model.parameters(), **kw)
    model.load_state_dict(state)
    _train(model, fx, spec, opt, 4, 90)                           # dirty the moments and step counts
    ptrs = [st["exp_avg"].data_ptr() for st in opt.state.values() if st]
    for want in fresh_runs:
        model.load_state_dict(state)
        opt.reset()
        assert all(float(st["step"]) == 0 and not st["exp_avg"].any() and not st["exp_avg_sq"].any()
                   for st in opt.state.values() if st)
        _train(model, fx, spec, opt, 3, 40)
        for k, v in model.state_dict().items():
            assert torch.equal(v, want[k]), k
    assert [st["exp_avg"].data_ptr() for st in opt.state.values() if st][:len(ptrs)] == ptrs
