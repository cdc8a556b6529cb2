"""step_graph.GraphedTrainStep: the whole train step (forward + bagging BCE + L2 + backward + FusedAdam) recorded as one
CUDA graph per (mask, batch shape) gives bit-identical parameters, optimizer state and losses to the eager step --
with dropout (seed read from device memory), across alternating masks, and when eager steps are mixed in (the device
step counters stay equal to the host's)."""
import importlib

import pytest
import torch

from tests.test_arena_gpu import DEV, _setup

pytestmark = pytest.mark.gpu
optim = importlib.import_module("aread-multi-domain-recommendation_b200.optim")
step_graph = importlib.import_module("aread-multi-domain-recommendation_b200.step_graph")


def _run(graphed, dropout, order, eager_at=()):
    fx, model, masks, batches = _setup(dropout=dropout)
    domains = [0, 1, 2, 3]
    for d in domains:
        model.domain_mask[d] = [m.to(DEV) for m in masks[d]]
    opt = optim.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    model.fold_regularization_into(opt)
    runner = step_graph.GraphedTrainStep(model, opt)
    losses = []
    for i, d in enumerate(order):
        torch.manual_seed(500 + i)
        x, y = batches[i % 4][0].to(DEV), batches[i % 4][1].to(DEV)
        if graphed and i not in eager_at:
            losses.append(float(runner(x, y, d)))
        else:
            losses.append(float(runner._eager(x, y, d)))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    opt_state = {i: (float(st["step"]), st["exp_avg"].clone(), st["exp_avg_sq"].clone())
                 for i, st in enumerate(opt.state.values()) if st}
    n_graphs = sum(e.graph is not None for e in runner.entries.values())
    return losses, state, opt_state, n_graphs


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_whole_step_graph_matches_eager(dropout):
    order = [0, 0, 0, 1, 1, 0, 1, 2, 2, 0, 2, 1, 3, 3, 3, 0]
    want = _run(False, dropout, order)
    got = _run(True, dropout, order, eager_at=(9,))
    assert got[3] == 4
    assert got[0] == want[0], "losses"
    for k in want[1]:
        assert torch.equal(got[1][k], want[1][k]), k
    assert got[2].keys() == want[2].keys()
    for i in want[2]:
        assert got[2][i][0] == want[2][i][0], f"step count of parameter {i}"
        assert torch.equal(got[2][i][1], want[2][i][1]) and torch.equal(got[2][i][2], want[2][i][2]), i


def test_requires_fused_adam_with_folded_regulariser():
    fx, model, masks, batches = _setup()
    with pytest.raises(TypeError):
        step_graph.GraphedTrainStep(model, torch.optim.Adam(model.parameters()))
    opt = optim.FusedAdam(model.parameters())
    with pytest.raises(RuntimeError):
        step_graph.GraphedTrainStep(model, opt)
