"""Parity of the CUDA AREAD module with the reference goldens and with the oracle on the same
seeded inputs.

The expert Linear layers run on the tensor cores in one of two operand precisions; everything else
is fp32.  Stated tolerances:

'bf16x3' (split operands, three passes, fp32-grade): against the fp32 reference goldens
  probabilities / logits rel 1e-4 (+ abs 1e-4 / 2e-4), loss rel 1e-4, every gradient tensor to a
  normalised error of 2e-2 (5e-2 for the gate parameters, whose gradients are small differences of
  large terms), post-Adam eval within 2e-3.

'bf16' (one pass, BASELINE.json configs[1] "bf16 experts"):
  * eval-mode probabilities / logits against the fp32 goldens: rel 1e-3 (+ abs 1e-3 / 2e-3);
  * train mode on these 37..64-row batches, where BatchNorm re-normalises the rounded activations:
    probabilities |d| <= 1e-2, loss rel 5e-3 against the fp32 goldens;
  * against the oracle evaluated with the SAME operand rounding (Spec.expert_operand_dtype=bf16):
    probabilities |d| <= 2e-3, loss rel 1e-3, and every gradient tensor closer to that oracle than
    max(2e-1 (gate parameters 4e-1), 1.5 x the effect the operand rounding itself has on that gradient);
  * |dAUC| < 1e-4 after 30 steps on identical weights.

Gate-mean side outputs (the HEMP thresholds compare them) are fp32 in both modes: round-off only."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import aread_torch as O
from oracle import synth
from tests._models import build_model
from tests._util import CASES, assert_close, family_errors, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

TOL = {
    "bf16x3": dict(prob=(1e-4, 1e-4), logit=(1e-4, 2e-4), train_prob=2e-4, loss=1e-4, grad=2e-2, grad_gate=5e-2,
                   after=2e-3),
    "bf16": dict(prob=(1e-3, 1e-3), logit=(1e-3, 2e-3), train_prob=1e-2, loss=5e-3, grad=None, grad_gate=None,
                 after=1e-2),
}
SAME_ROUNDING_PROB_ATOL = 2e-3             # bf16 mode vs the oracle with bf16 expert operands
# ... and its gradients, per tensor: normalised error below max(this, 1.5 x what the operand rounding itself does to
# that tensor).  Gate parameters (their gradients are small differences of large terms) get the wider base.
SAME_ROUNDING_GRAD, SAME_ROUNDING_GRAD_GATE = 2e-1, 4e-1
PRE_BN_BIAS = re.compile(r"\.layers\.(0|4|8)\.bias$")
GATE_PARAM = re.compile(r"^(tower_gates|mmoe_gates|group_embedding)\.")
PRECISIONS = ("bf16x3", "bf16")


def logit(p):
    p = p.double().clamp(1e-12, 1 - 1e-12)
    return torch.log(p) - torch.log1p(-p)


def check_probs(got, ref, what, tol):
    assert_close(got, ref, tol["prob"][0], tol["prob"][1], what)
    assert_close(logit(got.cpu()).float(), logit(ref).float(), tol["logit"][0], tol["logit"][1], what + " (logits)")


def check_grads(model, golden_grads, tol, what):
    """Gradients against the reference's, per parameter family (tests/_util.family_errors); pre-BatchNorm biases and
    other exactly-zero gradients must be ~0."""
    pairs = []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        comp = golden_grads[k]
        ref = comp["full"] if "full" in comp else comp["sample"]
        got = p.grad.detach().cpu().float()
        got = got if "full" in comp else got.reshape(-1)[::comp["stride"]][:ref.numel()]
        if PRE_BN_BIAS.search(k) or float(ref.abs().max()) < 1e-7:
            assert float(got.abs().max()) < 1e-4, f"{what} grad {k} should be ~0"
            continue
        pairs.append((k, got.reshape(-1), ref.float().reshape(-1)))
    if tol["grad"] is None:
        return
    for fam, (err, _) in family_errors(pairs).items():
        limit = tol["grad_gate"] if fam in ("tower_gates", "mmoe_gates") else tol["grad"]
        assert err < limit, f"{what} gradient of family {fam}: normalised error {err:.3e} (limit {limit})"


def grad_of(compact):
    return compact["full"] if "full" in compact else None


def _setup(name, dropout=0.0, precision="bf16"):
    fx = load_golden(name)
    spec = O.Spec(**fx["spec"])
    b0 = synth.random_batch(spec, fx["B"], seed=11, domain=fx["domain"], pad_id=fx["pad_id"])
    b1 = synth.random_batch(spec, fx["B"], seed=12, domain=fx["domain"], pad_id=fx["pad_id"])
    model = build_model(spec, DEV, dropout=dropout)
    model.expert_precision = precision
    return fx, spec, model, b0, b1


def to_dev(mask):
    return [m.clone().to(DEV) for m in mask]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASES)
def test_eval_modes_match_reference(name, precision):
    tol = TOL[precision]
    fx, spec, model, (x, _), _ = _setup(name, dropout=0.2, precision=precision)
    model.eval()
    ev = fx["eval"]
    xg = x.to(DEV)
    with torch.no_grad():
        y = model(xg, mode="wo_mask")
        assert tuple(y.shape) == tuple(ev["wo_mask"].shape)
        check_probs(y, ev["wo_mask"], "wo_mask", tol)
        for mk, m in fx["masks"].items():
            y = model(xg, mode="domain_with_mask", current_mask=to_dev(m))
            check_probs(y, ev[f"with_mask/{mk}"], f"with_mask/{mk}", tol)
            ys = model(xg, mode="domain_mask_bagging", current_mask=to_dev(m), tmp_memory_gate_value=True)
            assert tuple(ys.shape) == tuple(ev[f"bagging/{mk}"].shape)
            check_probs(ys, ev[f"bagging/{mk}"], f"bagging/{mk}", tol)
            for (l, t), ref in ev[f"gate_means/{mk}"].items():
                # HEMP thresholds compare these values: fp32 round-off only
                assert_close(model.tmp_tower_gate_values[l][t], ref, 1e-5, 1e-6, f"gate mean {l},{t}")
        y1 = model(xg[:1], mode="domain_with_mask", current_mask=to_dev(fx["masks"]["sparse"]))
        check_probs(y1, ev["with_mask/b1"], "batch of one (BatchNorm skipped)", tol)
        assert_close(model.get_regularization_loss(device=torch.device(DEV)), ev["reg"], 1e-5, 0, "reg")
        # domain_mask_final runs on full masks only (like the reference)
        model.domain_mask[fx["domain"]] = to_dev(fx["masks"]["full"])
        yf = model(xg, mode="domain_mask_final", domain_i=fx["domain"])
        assert tuple(yf.shape) == (fx["B"],) and bool(((yf > 0) & (yf < 1)).all())
    with pytest.raises(NameError):
        model(xg, mode="with_mask")


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mk", ["full", "sparse"])
def test_train_step_matches_reference(name, mk, precision):
    tol = TOL[precision]
    fx, spec, model, b0, b1 = _setup(name, precision=precision)
    tr = fx[f"train/{mk}"]
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    crit = torch.nn.BCELoss()
    mask = to_dev(fx["masks"][mk])
    for step in range(3):
        xb, yb = b0 if step % 2 == 0 else b1
        preds = model(xb.to(DEV), mode="domain_mask_bagging", current_mask=[m.clone() for m in mask],
                      tmp_memory_gate_value=True)
        tgt = yb.to(DEV).squeeze().float()
        data_loss = sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0]
        reg = model.get_regularization_loss(device=torch.device(DEV))
        loss = data_loss + reg
        model.zero_grad()
        loss.backward()
        if step == 0:
            assert_close(preds.detach(), tr["y_stack"], 0, tol["train_prob"], "y_stack")
            assert_close(data_loss, tr["data_loss"], tol["loss"], 1e-5, "data loss")
            assert_close(reg, tr["reg"], 1e-5, 0, "reg")
            for (l, t), ref in tr["gate_means"].items():
                assert_close(model.tmp_tower_gate_values[l][t], ref, 1e-5, 1e-6, f"gate mean {l},{t}")
            none_keys = sorted(k for k, p in model.named_parameters() if p.grad is None)
            assert none_keys == tr["grad_none"], "set of parameters without gradient"
            check_grads(model, tr["grads"], tol, "train")
        opt.step()
        assert_close(loss, tr[f"loss{step}"], tol["loss"] * (1 + 4 * step), 1e-5, f"loss{step}")
    model.eval()
    with torch.no_grad():
        y = model(b0[0].to(DEV), mode="domain_with_mask", current_mask=[m.clone() for m in mask])
    assert_close(y, tr["eval_after"], 0, tol["after"], "eval after 3 steps")
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked") and k in tr["state_after"]:
            assert int(v) == int(tr["state_after"][k]["full"]), k          # skipped towers are not tracked


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", CASES)
def test_wo_mask_train_matches_reference(name, precision):
    tol = TOL[precision]
    fx, spec, model, (x, y), _ = _setup(name, precision=precision)
    tr = fx["train/wo_mask"]
    model.train()
    dom = fx["domain"]
    pred = model(x.to(DEV), mode="wo_mask", domain_i=dom, memory_gate_value=True)
    loss = torch.nn.BCELoss()(pred.squeeze(), y.to(DEV).squeeze().float()) + \
        model.get_regularization_loss(device=torch.device(DEV))
    model.zero_grad()
    loss.backward()
    assert_close(pred.detach(), tr["y"], 0, tol["train_prob"], "y")
    assert_close(loss, tr["loss"], tol["loss"], 1e-5, "loss")
    for (l, t), ref in tr["recorded"].items():
        assert_close(model.domain_tower_gate_values[dom][l][t][0], ref, 1e-5, 1e-6, f"recorded gate {l},{t}")
    assert sorted(k for k, p in model.named_parameters() if p.grad is None) == tr["grad_none"]
    check_grads(model, tr["grads"], tol, "wo_mask")               # warm-up gradients (run.py:597-603)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mk", ["full", "sparse"])
def test_train_step_matches_oracle_with_same_operand_rounding(name, mk):
    """bf16 mode.  Same weights, inputs and mask; the oracle rounds the expert Linear operands to bf16
    and -- where the kernels store it that way -- the bias-free pre-activation, exactly as the kernels do, so what
    is left is accumulation order and the bf16 rounding of dy / dz in the backward.  Gradients that are small differences of large terms react strongly to ANY rounding, so
    each tensor is required to stay within 1.5 x the distance the operand rounding itself puts between
    the bf16 and the fp32 oracle (never tighter than 2e-1)."""
    fx, spec, model, (x, y), _ = _setup(name, precision="bf16")
    mask = fx["masks"][mk]
    grads = {}
    fused_bn = all(n % 64 == 0 for n in spec.expert_dims)      # the kernels then store the pre-activation as bf16
    for tag, dtype in (("bf16", torch.bfloat16), ("fp32", None)):
        sp = O.Spec(**fx["spec"], expert_operand_dtype=dtype,
                    expert_preact_dtype=dtype if fused_bn else None)
        sd = O.make_leaf_params(synth.deterministic_state(sp))
        out = O.forward(sd, sp, x, "domain_mask_bagging", mask, training=True)
        ref_loss = O.bagging_loss(out["y"], y) + O.reg_loss(sd, sp)
        ref_loss.backward()
        grads[tag] = {k: v.grad for k, v in sd.items() if v.requires_grad}
        if tag == "bf16":
            ref_y, ref_l = out["y"].detach(), ref_loss.detach()
    model.train()
    preds = model(x.to(DEV), mode="domain_mask_bagging", current_mask=to_dev(mask), tmp_memory_gate_value=True)
    tgt = y.to(DEV).squeeze().float()
    loss = sum(torch.nn.functional.binary_cross_entropy(p, tgt) for p in preds.unbind(0)) / preds.shape[0] + \
        model.get_regularization_loss(device=torch.device(DEV))
    model.zero_grad()
    loss.backward()
    assert_close(preds.detach(), ref_y, 0, SAME_ROUNDING_PROB_ATOL, "y_stack")
    assert_close(loss.detach(), ref_l, 1e-3, 1e-5, "loss")
    report, bad = [], []
    for k, p in model.named_parameters():
        ref = grads["bf16"].get(k)
        if p.grad is None:
            assert ref is None, k
            continue
        if PRE_BN_BIAS.search(k) or float(ref.abs().max()) < 1e-7:
            continue
        norm = float(ref.norm()) + 1e-12
        err = float((p.grad.cpu() - ref).norm()) / norm
        rounding_effect = float((ref - grads["fp32"][k]).norm()) / norm
        base = SAME_ROUNDING_GRAD_GATE if GATE_PARAM.search(k) else SAME_ROUNDING_GRAD
        report.append((err, rounding_effect, k))
        if not err < max(base, 1.5 * rounding_effect):
            bad.append(f"grad {k}: normalised error {err:.3e} (operand rounding effect {rounding_effect:.3e})")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_same_rounding.txt"), "a") as fh:
            for err, eff, k in sorted(report, reverse=True)[:12]:
                fh.write(f"{name} {mk} {k}: err {err:.3e} rounding effect {eff:.3e}\n")
    assert not bad, "; ".join(bad[:6])


def test_auc_after_training_matches_oracle():
    """30 train steps on the GPU and on the fp32 CPU oracle: the loss trajectories agree, and on
    IDENTICAL trained weights the two paths give |dAUC| < 1e-4 and logits within rel 1e-3 (the
    two independently trained weight sets differ more: Adam amplifies round-off, see
    tests/_util.assert_after_adam)."""
    from sklearn.metrics import roc_auc_score
    fx = load_golden("ali_small")
    spec = O.Spec(**fx["spec"])
    dom = fx["domain"]
    model = build_model(spec, DEV, dropout=0.0).train()
    model.expert_precision = "bf16"
    sd = O.make_leaf_params(synth.deterministic_state(spec))
    opt_g = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    opt_c = O.make_adam(sd)
    mask = fx["masks"]["sparse"]
    mask_g = to_dev(mask)
    crit = torch.nn.BCELoss()
    for step in range(30):
        x, y = synth.random_batch(spec, 256, seed=100 + step, domain=dom)
        loss_c, _ = O.train_step(sd, spec, x, y, mask, opt_c)
        preds = model(x.to(DEV), mode="domain_mask_bagging", current_mask=[m.clone() for m in mask_g])
        tgt = y.to(DEV).squeeze().float()
        loss_g = sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0] + \
            model.get_regularization_loss(device=torch.device(DEV))
        model.zero_grad()
        loss_g.backward()
        opt_g.step()
        assert abs(float(loss_g.detach()) - float(loss_c)) < 5e-3 * abs(float(loss_c)) + 1e-4, step
    x, y = synth.random_batch(spec, 4096, seed=999, domain=dom)
    labels = y.numpy().reshape(-1)
    model.eval()
    with torch.no_grad():
        p_g = model(x.to(DEV), mode="domain_with_mask", current_mask=[m.clone() for m in mask_g]).cpu()
        p_c = O.forward(sd, spec, x, "domain_with_mask", mask)["y"]
        # the GPU-trained weights evaluated by the oracle
        sd_g = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        p_same = O.forward(sd_g, spec, x, "domain_with_mask", mask)["y"]
    assert abs(roc_auc_score(labels, p_g.numpy()) - roc_auc_score(labels, p_same.numpy())) < 1e-4
    check_probs(p_g, p_same, "eval on identical trained weights", TOL["bf16"])
    assert abs(roc_auc_score(labels, p_g.numpy()) - roc_auc_score(labels, p_c.numpy())) < 5e-3
