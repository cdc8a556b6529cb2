"""Host-side HEMP bookkeeping of this repository's AREAD against masks recorded from the
reference with fixed seeds (tests/golden/hemp.pt).  CPU only: nothing here launches a kernel."""
import copy

import numpy as np
import torch

from oracle import aread_torch as O
from tests._models import build_model
from tests._util import assert_close, load_golden


def _same(a, b):
    return len(a) == len(b) and all(torch.equal(torch.as_tensor(x).bool().cpu(), torch.as_tensor(y).bool().cpu())
                                    for x, y in zip(a, b))


def _model():
    fx = load_golden("hemp")
    spec = O.Spec(**fx["spec"])
    return fx, spec, build_model(spec, "cpu").eval()


def test_validate_mask_numpy_and_tensor():
    fx, spec, model = _model()
    for i, (raw, valid) in enumerate(zip(fx["validate"]["raw"], fx["validate"]["valid"])):
        if i % 2 == 0:
            arg = [r.numpy().copy() for r in raw]
        else:
            arg = [r.clone() for r in raw]
        got = model.validate_mask(arg)
        assert got is arg                      # in-place contract
        assert _same(got, valid), f"mask {i}"


def test_generate_rand_sequence():
    fx, spec, model = _model()
    np.random.seed(17)
    for p, ref in zip((0.7, 0.4, 0.2, 0.1), fx["rand"]):
        got = model.generate_mask("rand", 0, init_active_percent=p)
        assert all(t.dtype == torch.bool for t in got)
        assert _same(got, ref), p


def _load_recorded(model, fx, spec):
    rec = model._empty_gate_log()
    for (l, t), vs in fx["recorded_d4"].items():
        rec[l][t] = [v.clone() for v in vs]
    return rec


def test_gate_driven_generation():
    fx, spec, model = _model()
    for gm in ("max_gate", "mask_max_gate", "max_gate_norm_rand", "mask_norm_rand"):
        model.domain_tower_gate_values[4] = _load_recorded(model, fx, spec)
        model.gate_value_threshold[4] = None
        model.domain_mask[4] = [t.clone() for t in fx["rand"][1]] if gm == "mask_norm_rand" else None
        np.random.seed(23)
        torch.manual_seed(29)
        got = model.generate_mask(gm, 4, init_active_percent=0.5, random_modify_sigma=0.2)
        assert _same(got, fx["generate"][gm]), gm
        if gm == "max_gate":
            for a, b in zip(model.domain_tower_gate_values[4], fx["mean_values_d4"]):
                assert_close(a, b, 1e-6, 1e-8, "mean gate values")
            assert_close(torch.as_tensor(model.gate_value_threshold[4]), fx["threshold_d4"], 1e-6, 0, "threshold")
    model.domain_tower_gate_values[4] = _load_recorded(model, fx, spec)
    model.domain_mask[4] = [t.clone() for t in fx["generate"]["max_gate"]]
    np.random.seed(31)
    torch.manual_seed(37)
    got = model.generate_mask("mask_max_gate", 4, init_active_percent=0.3, random_modify_sigma=0.2)
    assert _same(got, fx["generate"]["mask_max_gate/steady"])


def test_prune_single_mask():
    fx, spec, model = _model()
    for (l, t), v in fx["prune_in_gates"].items():
        model.tmp_tower_gate_values[l][t] = v.clone()
    mask = [t.clone() for t in fx["prune_in_mask"]]
    got = model.prun_single_mask(4, mask, prun_ratio=0.25)
    assert _same(got, fx["prune_out_mask"])
    assert all(v is None for lvl in model.tmp_tower_gate_values for v in lvl)


def test_update_all_mask_picks_lowest_loss():
    fx, spec, model = _model()
    model.reset_for_mask_update()
    for d in range(spec.n_domain):
        for z in range(3):
            model.candidate_domain_mask[d].append([t.clone() for t in fx["update_all"]["candidates"][d][z]])
            for s in range(2):
                model.add_eval_loss(float(((d * 7 + z * 3 + s) % 5) * 0.1 + 0.3), d, z)
    model.update_all_mask(regroup_times=1)
    for d in range(spec.n_domain):
        assert _same(model.domain_mask[d], fx["update_all"]["chosen"][d]), d
    assert abs(float(model.count_current_active_ratio()) - fx["update_all"]["active_ratio"]) < 1e-12


def test_state_dict_layout_and_rollback():
    from oracle import synth
    fx, spec, model = _model()
    ref_shapes = synth.state_shapes(spec, with_attention=True)
    sd = model.state_dict()
    assert set(sd) == set(ref_shapes)
    assert all(tuple(sd[k].shape) == tuple(ref_shapes[k]) for k in sd)
    model.save_model_state()
    kept = set(model.model_state)
    assert all(k.split(".")[0] in ("cn", "towers", "tower_gates", "towers_linear", "embedding", "linear") for k in kept)
    assert not any(k.startswith(("mmoe_", "group_embedding", "final_gate")) for k in kept)
    before = copy.deepcopy({k: v.clone() for k, v in sd.items()})
    with torch.no_grad():
        for p in model.parameters():
            p.add_(1.0)
    model.load_model_state()
    after = model.state_dict()
    for k in before:
        if k in kept:
            assert torch.equal(after[k], before[k]), k
        elif after[k].is_floating_point() and not k.endswith(("running_mean", "running_var")):
            assert not torch.equal(after[k], before[k]), k
    # reg groups: table, linear.fc.weight, experts (incl. BN gamma), towers (incl. BN gamma), cn.w
    sizes = [sum((w[1] if isinstance(w, tuple) else w).numel() for w in ws) for ws, _, _ in model.regularization_weight]
    E = spec.E
    assert sizes[0] == spec.n_rows * spec.embed_dim and sizes[1] == E and sizes[4] == spec.n_cross_layers * E
