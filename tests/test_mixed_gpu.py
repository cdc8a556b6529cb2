"""Mixed-domain batches (csrc/mixed.cu, mixed_ops.py): one call evaluates rows of >= 30 domains, each under its own
HEMP mask, and must equal -- row for row -- what the reference semantics give when the model is called once per
domain (run.py:719-727, model/aread.py:224-234): checked against this module's own per-domain path (fp32 round-off)
and against oracle/aread_torch.py on the CPU (bf16 experts: logits |d| <= 1e-3 |z| + 2e-3).  Row order must not
matter (sorted by domain or shuffled: bit-identical per row), pruned towers must not leak into the result, and the
per-domain gate means of a mixed 'wo_mask' batch equal the boolean-index means of aread.py:187-200."""
import importlib

import numpy as np
import pytest
import torch

from oracle import aread_torch as O
from oracle import synth
from tests._models import build_model
from tests._util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
mixed_ops = importlib.import_module("aread-multi-domain-recommendation_b200.mixed_ops")


def _model_with_masks(precision="bf16", seed=5):
    fx = load_golden("ali_small")
    spec = O.Spec(**fx["spec"])
    model = build_model(spec, DEV, dropout=0.2)
    model.expert_precision = precision
    np.random.seed(seed)
    for d in range(spec.n_domain):
        p = (0.15, 0.3, 0.5, 0.8, 1.0)[d % 5]
        model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=p) if p < 1 else \
            [m.to(DEV) for m in synth.full_mask(spec)]
    return fx, spec, model.eval()


def _mixed_batch(spec, B, seed, sort):
    rng = np.random.RandomState(seed)
    x, y = synth.random_batch(spec, B, seed=seed)
    sizes = rng.zipf(1.3, size=B) % spec.n_domain                     # skewed domain sizes, some domains tiny
    x[:, spec.domain_idx] = torch.from_numpy(sizes.astype(np.int32))
    if sort:
        order = torch.argsort(x[:, spec.domain_idx].long(), stable=True)
        x, y = x[order], y[order]
    return x, y


def _logit(p):
    p = p.double().clamp(1e-12, 1 - 1e-12)
    return torch.log(p) - torch.log1p(-p)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_mixed_eval_equals_per_domain_calls(precision):
    fx, spec, model = _model_with_masks(precision)
    x, _ = _mixed_batch(spec, 3000, seed=1, sort=True)
    xg = x.to(DEV)
    dom = x[:, spec.domain_idx].long()
    assert len(torch.unique(dom)) >= 25
    y, y_stack = model.forward_mixed(xg, return_stack=True)
    y, y_stack = y.cpu(), y_stack.cpu()
    sd = synth.deterministic_state(spec)
    for d in torch.unique(dom).tolist():
        rows = torch.nonzero(dom == d).squeeze(1)
        with torch.no_grad():
            want = model(xg[rows.to(DEV)], mode="domain_with_mask", domain_i=d).cpu()
            stack = model(xg[rows.to(DEV)], mode="domain_mask_bagging", domain_i=d).cpu()
        # same kernels up to the HEI levels; those differ in summation order only
        def same(a, b):      # fp32 summation order (BatchNorm folded into the weights vs applied afterwards)
            return bool(((_logit(a) - _logit(b)).abs() <= 2e-4 * _logit(b).abs() + 2e-4).all())
        assert same(y[rows], want), f"domain {d}"
        act = model.mask_info(model.domain_mask[d]).active_idx[-1]
        got_stack = y_stack[:, rows]
        assert same(got_stack[act], stack), f"domain {d} (heads)"
        off = [t for t in range(spec.n_tower[-1]) if t not in act]
        assert not got_stack[off].any(), "pruned heads report 0"
        # ... and the reference arithmetic itself (oracle, fp32 on the CPU)
        if len(rows) > 1 and d % 3 == 0:
            mask = [m.cpu() for m in model.domain_mask[d]]
            ref = O.forward(sd, spec, x[rows], "domain_with_mask", mask)["y"]
            err = (_logit(y[rows]) - _logit(ref)).abs()
            tol = (1e-3 if precision == "bf16" else 1e-4) * _logit(ref).abs() + (2e-3 if precision == "bf16" else 2e-4)
            assert bool((err <= tol).all()), f"domain {d} vs oracle: {float(err.max()):.3e}"


def test_row_order_does_not_matter_and_single_rows():
    fx, spec, model = _model_with_masks()
    x, _ = _mixed_batch(spec, 2000, seed=2, sort=True)
    y_sorted = model.forward_mixed(x.to(DEV)).cpu()
    perm = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(0))
    y_perm = model.forward_mixed(x[perm].to(DEV)).cpu()
    assert torch.equal(y_perm, y_sorted[perm]), "per-row results must not depend on the neighbours in the tile"
    # ragged tail / tiny batches / an id outside the domain range
    for n in (1, 2, 33):
        got = model.forward_mixed(x[:n].to(DEV)).cpu()
        if n > 1:                                  # a batch of one skips BatchNorm (layer.py:226): different function
            assert float((_logit(got) - _logit(y_sorted[:n])).abs().max()) <= 1e-5
        else:
            want = model(x[:1].to(DEV), mode="domain_with_mask", domain_i=int(x[0, spec.domain_idx]))
            assert float((_logit(got) - _logit(want.cpu())).abs().max()) <= 2e-4


def test_train_mode_is_refused():
    fx, spec, model = _model_with_masks()
    model.train()
    with pytest.raises(RuntimeError):
        model.forward_mixed(torch.zeros(4, len(spec.one_hot_field_dims), dtype=torch.int32, device=DEV))


def test_domain_means_kernel_and_mixed_gate_recording():
    fx, spec, model = _model_with_masks()
    x, _ = _mixed_batch(spec, 5000, seed=3, sort=False)
    xg = x.to(DEV)
    vals = torch.randn(5000, 72, device=DEV)
    mean, count = mixed_ops.domain_means(vals, xg, spec.domain_idx, spec.n_domain)
    dom = xg[:, spec.domain_idx].long()
    for d in range(spec.n_domain):
        sel = dom == d
        assert int(count[d]) == int(sel.sum())
        if sel.any():
            torch.testing.assert_close(mean[d], vals[sel].mean(dim=0), rtol=1e-5, atol=1e-6)
    # 'wo_mask' with memory_gate_value on a mixed batch (domain_i=None): aread.py:187-200
    model.reset_for_mask_update()
    with torch.no_grad():
        model(xg, mode="wo_mask", memory_gate_value=True)
        sd = synth.deterministic_state(spec)
        ref = O.forward(sd, spec, x, "wo_mask")
    for d in range(spec.n_domain):
        sel = (x[:, spec.domain_idx] == d)
        for l in range(1, spec.n_level):
            for t in range(spec.n_tower[l]):
                got = model.domain_tower_gate_values[d][l][t]
                assert len(got) == 1
                if sel.any():
                    want = ref["gates"][(l, t)][sel].mean(dim=0)
                    torch.testing.assert_close(got[0].cpu(), want, rtol=1e-4, atol=1e-6)
                else:
                    assert bool(torch.isnan(got[0]).all())
