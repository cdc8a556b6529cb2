#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
modules imported from /root/reference (build container only; the GPU box never runs this).

    python tests/golden/make_golden.py

Model states and inputs are pure functions of (key, shape, seed) (oracle/synth.py), so a
fixture only stores the reference's *outputs*.  Large tensors are stored as a strided sample
plus float64 sum / sum of squares.
"""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AREAD_REF", "/root/reference")
sys.path.insert(0, ROOT)

from oracle.aread_torch import Spec          # noqa: E402
from oracle import synth                     # noqa: E402

FULL_LIMIT = 1024
SAMPLE = 512


def compact(t):
    t = t.detach().cpu()
    if t.numel() <= FULL_LIMIT:
        return {"full": t.clone()}
    flat = t.reshape(-1)
    stride = max(1, flat.numel() // SAMPLE)
    d = flat.double()
    return {"sample": flat[::stride][:SAMPLE].clone(), "stride": stride, "shape": tuple(t.shape),
            "sum": float(d.sum()), "sumsq": float((d * d).sum())}


def load_reference():
    sys.path.insert(0, REF)
    import config as refcfg                 # noqa
    from model.aread import AREAD           # noqa
    return refcfg, AREAD


CASES = {
    # AliCCP-shaped: 23 one-hot fields, default expert/tower sizes, small vocabularies
    "ali_small": dict(
        spec=dict(one_hot_field_dims=[400, 95, 14, 3, 8, 4, 4, 3, 5, 300, 30, 500, 200, 250, 60, 260, 120, 90,
                                      40, 230, 110, 70, 4],
                  embed_dim=32, n_domain=30, domain_idx=10, itemid_idx=9),
        B=64, domain=7, pad_id=None),
    # Amazon-shaped: 7 one-hot + 2 x 5 multi-hot columns, mean pooled, padding id == item vocabulary
    # (aliases the first row of the next field)
    "amz_small": dict(
        spec=dict(one_hot_field_dims=[500, 7, 25, 45, 11, 300, 10], embed_dim=32,
                  multi_hot_flag=[False] * 7 + [True] * 10, itemid_idx=0, seq_maxlen=5, method="mean",
                  n_domain=25, domain_idx=2),
        B=48, domain=3, pad_id=500),
    # generality: odd sizes everywhere
    "tiny": dict(
        spec=dict(one_hot_field_dims=[37, 11, 6, 19, 5], embed_dim=8, n_domain=6, domain_idx=2, itemid_idx=0,
                  n_tower=(2, 4, 8), expert_dims=(32, 16, 8), tower_dims=((8, 8), (8, 4), (4, 4)), n_expert=3,
                  n_cross_layers=2),
        B=37, domain=1, pad_id=None),
}


def build_reference(refcfg, AREAD, spec, dropout):
    cfg = types.SimpleNamespace(**{k: v for k, v in vars(refcfg).items() if not k.startswith("__")})
    cfg.dataset_name = "synth"
    cfg.domain_size = {"synth": [100] * spec.n_domain}
    cfg.n_cross_layers = spec.n_cross_layers
    cfg.mmoe_n_expert = spec.n_expert
    cfg.use_dcn, cfg.use_atten = True, True
    mh = {"multi_hot_flag": list(spec.flag), "itemid_idx": spec.itemid_idx,
          "seq_maxlen": spec.seq_maxlen, "method": spec.method}
    model = AREAD(np.asarray(spec.one_hot_field_dims), spec.embed_dim, mh, n_tower=tuple(spec.n_tower),
                  n_domain=spec.n_domain, base_model="mmoe", expert_dims=tuple(spec.expert_dims),
                  tower_dims=tuple(tuple(t) for t in spec.tower_dims), domain_idx=spec.domain_idx,
                  device=torch.device("cpu"), dropout=dropout, config=cfg)
    model.reset_for_mask_update()
    sd = synth.deterministic_state(spec, with_attention=True)
    ref_sd = model.state_dict()
    assert set(ref_sd) == set(sd), (set(ref_sd) ^ set(sd))
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    return model


def sparse_mask(model, seed, p):
    np.random.seed(seed)
    return model.generate_mask("rand", 0, init_active_percent=p)


def gate_means_of(model, spec):
    out = {}
    for l in range(1, spec.n_level):
        for t in range(spec.n_tower[l]):
            v = model.tmp_tower_gate_values[l][t]
            out[(l, t)] = None if v is None else v.clone()
    return out


def run_case(name, case, refcfg, AREAD):
    spec = Spec(**case["spec"])
    B, dom = case["B"], case["domain"]
    x, y = synth.random_batch(spec, B, seed=11, domain=dom, pad_id=case["pad_id"])
    x2, y2 = synth.random_batch(spec, B, seed=12, domain=dom, pad_id=case["pad_id"])
    fx = {"name": name, "spec": case["spec"], "B": B, "domain": dom, "pad_id": case["pad_id"]}

    # ---- eval-mode forwards (running-stat BN, no dropout)
    model = build_reference(refcfg, AREAD, spec, dropout=0.2).eval()
    masks = {"full": synth.full_mask(spec), "sparse": sparse_mask(model, 5, 0.35)}
    fx["masks"] = {k: [m.clone() for m in v] for k, v in masks.items()}
    ev = {}
    with torch.no_grad():
        ev["embed"] = model.embedding(x).clone()
        ev["wo_mask"] = model(x, mode="wo_mask").clone()
        for mk, m in masks.items():
            ev[f"with_mask/{mk}"] = model(x, mode="domain_with_mask", current_mask=[t.clone() for t in m]).clone()
            ys = model(x, mode="domain_mask_bagging", current_mask=[t.clone() for t in m],
                       tmp_memory_gate_value=True)
            ev[f"bagging/{mk}"] = ys.clone()
            ev[f"gate_means/{mk}"] = gate_means_of(model, spec)
        # batch of one row: BatchNorm is skipped (layer.py:226)
        ev["with_mask/b1"] = model(x[:1], mode="domain_with_mask",
                                   current_mask=[t.clone() for t in masks["sparse"]]).clone()
        ev["reg"] = model.get_regularization_loss(device=torch.device("cpu")).clone()
    fx["eval"] = ev

    # ---- train-mode (batch-stat BN, dropout 0): grads of step 0, then 3 Adam steps
    for mk in ("full", "sparse"):
        model = build_reference(refcfg, AREAD, spec, dropout=0.0).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        crit = torch.nn.BCELoss()
        tr = {}
        for step in range(3):
            xb, yb = (x, y) if step % 2 == 0 else (x2, y2)
            preds = model(xb, mode="domain_mask_bagging", current_mask=[t.clone() for t in masks[mk]],
                          tmp_memory_gate_value=True)
            tgt = yb.squeeze().float()
            losses = [crit(p, tgt) for p in preds.unbind(dim=0)]
            data_loss = sum(losses) / preds.shape[0]
            reg = model.get_regularization_loss(device=torch.device("cpu"))
            loss = data_loss + reg
            model.zero_grad()
            loss.backward()
            if step == 0:
                tr["y_stack"] = preds.detach().clone()
                tr["data_loss"] = data_loss.detach().clone()
                tr["reg"] = reg.detach().clone()
                tr["gate_means"] = gate_means_of(model, spec)
                tr["grad_none"] = sorted(k for k, p in model.named_parameters() if p.grad is None)
                tr["grads"] = {k: compact(p.grad) for k, p in model.named_parameters() if p.grad is not None}
            opt.step()
            tr[f"loss{step}"] = loss.detach().clone()
        tr["state_after"] = {k: compact(v) for k, v in model.state_dict().items()
                             if not k.startswith(("atten_", "self_attns", "V_res", "final_gate"))}
        model.eval()
        with torch.no_grad():
            tr["eval_after"] = model(x, mode="domain_with_mask",
                                     current_mask=[t.clone() for t in masks[mk]]).clone()
        fx[f"train/{mk}"] = tr

    # ---- wo_mask train step (warm-up path, run.py:597-603) incl. recorded gate values
    model = build_reference(refcfg, AREAD, spec, dropout=0.0).train()
    pred = model(x, mode="wo_mask", domain_i=dom, memory_gate_value=True)
    loss = torch.nn.BCELoss()(pred.squeeze(), y.squeeze().float()) + model.get_regularization_loss(torch.device("cpu"))
    model.zero_grad()
    loss.backward()
    fx["train/wo_mask"] = {
        "y": pred.detach().clone(), "loss": loss.detach().clone(),
        "recorded": {(l, t): model.domain_tower_gate_values[dom][l][t][0].clone()
                     for l in range(1, spec.n_level) for t in range(spec.n_tower[l])},
        "grad_none": sorted(k for k, p in model.named_parameters() if p.grad is None),
        "grads": {k: compact(p.grad) for k, p in model.named_parameters() if p.grad is not None},
    }
    torch.save(fx, os.path.join(HERE, f"{name}.pt"))
    print("wrote", name, os.path.getsize(os.path.join(HERE, f"{name}.pt")) // 1024, "KiB")


def run_hemp(refcfg, AREAD):
    """Host-side HEMP bookkeeping (aread.py:330-605): masks produced from fixed seeds."""
    spec = Spec(**CASES["ali_small"]["spec"])
    model = build_reference(refcfg, AREAD, spec, dropout=0.0).eval()
    out = {"spec": CASES["ali_small"]["spec"]}
    # validate_mask on raw random masks
    rng = np.random.RandomState(3)
    raws, valids = [], []
    nt = spec.n_tower
    for i in range(24):
        p = [0.15, 0.3, 0.5, 0.8][i % 4]
        raw = [rng.rand(1, nt[0]) < p] + [rng.rand(nt[l - 1], nt[l]) < p for l in range(1, 3)] + [rng.rand(nt[-1], 1) < p]
        raws.append([torch.tensor(r) for r in raw])
        if i % 2 == 0:
            v = model.validate_mask([r.copy() for r in raw])
            valids.append([torch.tensor(a) for a in v])
        else:
            v = model.validate_mask([torch.tensor(r) for r in raw])
            valids.append([a.clone() for a in v])
    out["validate"] = {"raw": raws, "valid": valids}
    # generate_mask('rand')
    np.random.seed(17)
    out["rand"] = [[t.clone() for t in model.generate_mask("rand", 0, init_active_percent=p)]
                   for p in (0.7, 0.4, 0.2, 0.1)]
    # gate-value driven modes: feed recorded gate values, then generate
    x, _ = synth.random_batch(spec, 32, seed=21, domain=4)
    with torch.no_grad():
        for _ in range(3):
            model(x, mode="wo_mask", domain_i=4, memory_gate_value=True)
    rec = copy.deepcopy(model.domain_tower_gate_values[4])
    out["recorded_d4"] = {(l, t): [v.clone() for v in rec[l][t]] for l in range(1, 3) for t in range(nt[l])}
    gen = {}
    for gm in ("max_gate", "mask_max_gate", "max_gate_norm_rand", "mask_norm_rand"):
        model.domain_tower_gate_values[4] = copy.deepcopy(rec)
        model.gate_value_threshold[4] = None
        model.domain_mask[4] = [t.clone() for t in out["rand"][1]] if gm in ("mask_norm_rand",) else None
        np.random.seed(23)
        torch.manual_seed(29)
        m = model.generate_mask(gm, 4, init_active_percent=0.5, random_modify_sigma=0.2)
        gen[gm] = [t.clone() for t in m]
        if gm == "max_gate":
            out["mean_values_d4"] = [t.clone() for t in model.domain_tower_gate_values[4]]
            out["threshold_d4"] = torch.as_tensor(model.gate_value_threshold[4]).clone()
    # second call of mask_max_gate with an existing domain mask (the steady-state path, run.py:627)
    model.domain_tower_gate_values[4] = copy.deepcopy(rec)
    model.domain_mask[4] = [t.clone() for t in gen["max_gate"]]
    np.random.seed(31)
    torch.manual_seed(37)
    gen["mask_max_gate/steady"] = [t.clone() for t in
                                   model.generate_mask("mask_max_gate", 4, init_active_percent=0.3,
                                                       random_modify_sigma=0.2)]
    out["generate"] = gen
    # prun_single_mask: run a masked forward to fill tmp gate values, then prune
    m0 = [t.clone() for t in out["rand"][0]]
    with torch.no_grad():
        model(x, mode="domain_mask_bagging", current_mask=m0, tmp_memory_gate_value=True)
    out["prune_in_gates"] = {(l, t): model.tmp_tower_gate_values[l][t].clone()
                             for l in range(1, 3) for t in range(nt[l])}
    out["prune_in_mask"] = [t.clone() for t in m0]
    pruned = model.prun_single_mask(4, m0, prun_ratio=0.25)
    out["prune_out_mask"] = [t.clone() for t in pruned]
    # update_all_mask: pick the candidate with the lowest mean eval loss
    model.reset_for_mask_update()
    np.random.seed(41)
    cands = []
    for d in range(spec.n_domain):
        for z in range(3):
            model.candidate_domain_mask[d].append(model.generate_mask("rand", d, init_active_percent=0.5))
            for s in range(2):
                model.add_eval_loss(float(((d * 7 + z * 3 + s) % 5) * 0.1 + 0.3), d, z)
        cands.append([[t.clone() for t in m] for m in model.candidate_domain_mask[d]])
    model.update_all_mask(regroup_times=1)
    out["update_all"] = {"candidates": cands, "chosen": [[t.clone() for t in m] for m in model.domain_mask],
                         "active_ratio": float(model.count_current_active_ratio())}
    torch.save(out, os.path.join(HERE, "hemp.pt"))
    print("wrote hemp", os.path.getsize(os.path.join(HERE, "hemp.pt")) // 1024, "KiB")


def main():
    torch.set_num_threads(4)
    refcfg, AREAD = load_reference()
    for name, case in CASES.items():
        run_case(name, case, refcfg, AREAD)
    run_hemp(refcfg, AREAD)


if __name__ == "__main__":
    main()
