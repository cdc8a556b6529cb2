"""The reference trainer's call sequence (run.py:578-686: warm-up -> regroup with save/load_model_state,
generate_mask('mask_max_gate'), bagging steps + prun_single_mask, no_grad scoring in train mode, update_all_mask ->
training under the selected masks) replayed on the CUDA module and compared with the record the UNMODIFIED
reference produced for the same seeded sequence on CPU (tests/golden/trainer_seq.pt, written by
tests/golden/make_trainer_golden.py).

Experts run in 'bf16x3' (fp32-grade) and the sequence uses a small learning rate (see tests/_trainer_sequence.py
for why) so that the HEMP decisions -- quantile thresholds on recorded gate means -- see the reference's numbers to
round-off: candidate masks, pruned masks and the finally selected masks must be IDENTICAL; losses agree to rel
2e-3."""
import importlib

import pytest
import torch

from oracle import aread_torch as O
from tests import _trainer_sequence as T
from tests._models import build_model
from tests._util import assert_after_adam, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
fused = importlib.import_module("aread-multi-domain-recommendation_b200.fused")


def _close(a, b, rtol, what):
    assert len(a) == len(b), what
    for i, (u, v) in enumerate(zip(a, b)):
        assert abs(u - v) <= rtol * abs(v) + 1e-5, f"{what}[{i}]: {u} vs {v}"


@pytest.mark.parametrize("graphs", [False, True])
def test_train_aread_sequence_matches_reference(monkeypatch, graphs):
    gold = load_golden("trainer_seq")
    assert gold["spec"] == T.SPEC and gold["cfg"] == T.SEQ
    spec = O.Spec(**T.SPEC)
    monkeypatch.setattr(fused, "USE_GRAPHS", graphs)
    model = build_model(spec, DEV, dropout=0.0)
    model.expert_precision = "bf16x3"
    got = T.run_sequence(model, torch.device(DEV), spec)
    _close(got["warm_up"], gold["warm_up"], 2e-3, "warm-up loss")
    # parameters after the warm-up (8 Adam steps): Adam moves every element by about lr per step whatever the size of
    # its gradient, so agreement is required to a fraction of the distance travelled (tests/_util.assert_after_adam)
    worst = sorted(((float((got["after_warm_up"][k] - v).abs().max()), k) for k, v in gold["after_warm_up"].items()),
                   reverse=True)[:5]
    for k, v in gold["after_warm_up"].items():
        assert_after_adam(got["after_warm_up"][k], {"full": v}, T.SEQ["warm_up"], T.LR, f"{k} after warm-up (worst: {worst})")
    # the quantities HEMP thresholds (batch means of the masked gate softmax) before every prune, then the prune itself
    for i, (g, r) in enumerate(zip(got["gate_log"], gold["gate_log"])):
        for l, (a, b) in enumerate(zip(g, r)):
            assert float((a - b).abs().max()) <= 2e-4, f"gate means before prune {i}, level {l + 1}: {(a - b).abs().max()}"
        assert got["prune_log"][i] == gold["prune_log"][i], \
            f"prun_single_mask #{i}: gate means {[x.tolist() for x in g]} vs reference {[x.tolist() for x in r]}"
    for g, r in zip(got["candidates"], gold["candidates"]):
        assert (g["d"], g["z"]) == (r["d"], r["z"])
        assert g["generated"] == r["generated"], f"generate_mask('mask_max_gate') of domain {g['d']}, candidate {g['z']}"
        assert g["pruned"] == r["pruned"], f"prun_single_mask chain of domain {g['d']}, candidate {g['z']}"
    for g, r in zip(got["update"], gold["update"]):
        _close(g, r, 2e-3, "candidate update loss")
    for g, r in zip(got["eval_loss"], gold["eval_loss"]):
        _close(g, r, 2e-3, "candidate eval loss")
    assert got["selected"] == gold["selected"], "update_all_mask selection"
    _close(got["post"], gold["post"], 5e-3, "loss under the selected masks")
    assert got["recorded_gate_counts"] == gold["recorded_gate_counts"]
