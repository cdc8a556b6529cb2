"""The call sequence of the reference trainer's `Run.train_aread` (run.py:578-686), restated as a plain function
over any AREAD-shaped module (TEST INFRASTRUCTURE).

    warm-up ('wo_mask', memory_gate_value=True)                                   run.py:588-606
    regroup: save_model_state -> per domain, per candidate:
        generate_mask('mask_max_gate') -> load_model_state -> fresh Adam          run.py:614-633
        regroup_update_step x [bagging step under the candidate + prun_single_mask]   run.py:634-648
        candidate_domain_mask[d].append ; regroup_eval_step x no_grad 'domain_with_mask' (train mode) + add_eval_loss
    update_all_mask -> reset_for_mask_update -> load_model_state                  run.py:659-661
    bagging steps under the selected masks                                        run.py:663-682

The same function drives the unmodified reference on CPU (tests/golden/make_trainer_golden.py, build container) and
this repository's CUDA module (tests/test_trainer_sequence_gpu.py); what it returns is what the two are compared on.
Data, model state and every random draw are seeded; `torch.rand(..., device=...)` inside `generate_mask` is routed
through the CPU generator so that both sides draw the same flips.
"""
import contextlib

import numpy as np
import torch

from oracle import synth

# Adam moves every element by about lr per step whatever the size of its gradient, so after a few steps two fp32
# implementations that differ in summation order only are up to 2 * steps * lr apart in parameters whose gradient is
# small and noisy (the gate Linears: measured on the CPU oracle, rounding the expert operands to bf16 alone moves the
# level-1 gate means by 1.5e-2 after the 8 warm-up steps at lr = 1e-3).  HEMP then thresholds exactly those values, so
# at the trainer's lr the chosen masks legitimately differ between any two implementations.  The sequence is therefore
# replayed with a learning rate small enough (1e-6) that the decisions are determined by the common starting point:
# every call, every piece of bookkeeping and every mask must then agree with the reference exactly.
LR, UPDATE_LR, WD = 1e-6, 1e-6, 1e-8

SEQ = dict(n_domain=4, B=64, warm_up=8, candidates=2, update_steps=5, eval_steps=5, post_steps=6,
           init_active_percent=0.7, random_modify_sigma=0.2, seed=2000)

SPEC = dict(one_hot_field_dims=[300, 40, 14, 3, 8, 4, 220, 4, 150, 60, 90, 5], embed_dim=32, n_domain=4,
            domain_idx=7, itemid_idx=6)


@contextlib.contextmanager
def cpu_rand():
    """torch.rand on the CPU generator, moved to the requested device afterwards."""
    real = torch.rand

    def rand(*size, device=None, **kw):
        out = real(*size, **kw)
        return out if device is None else out.to(device)
    torch.rand = rand
    try:
        yield
    finally:
        torch.rand = real


def batches(spec, cfg):
    """Per-domain batch streams ('train' and 'aug_train' of run.py:551-575) as seeded generators."""
    counters = {}

    def get(d, mode="train"):
        i = counters.get((d, mode), 0)
        counters[(d, mode)] = i + 1
        seed = cfg["seed"] + 7919 * d + 104729 * (mode == "aug_train") + i
        return synth.random_batch(spec, cfg["B"], seed=seed, domain=d)
    return get


def mask_to_lists(mask):
    return [m.detach().cpu().numpy().astype(bool).tolist() for m in mask]


def run_sequence(model, device, spec, cfg=SEQ, loss_fn=None):
    """Returns a dict of everything observable: losses of every phase, candidate / selected masks, eval losses."""
    crit = torch.nn.BCELoss()
    get = batches(spec, cfg)
    n_domain = cfg["n_domain"]
    out = {"warm_up": [], "candidates": [], "eval_loss": [], "selected": [], "post": [], "update": [], "gate_log": [],
           "prune_log": []}
    torch.manual_seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=LR, betas=(0.9, 0.99), eps=1e-8, weight_decay=WD)

    def to_dev(x, y):
        return x.to(device), y.to(device)

    def bagging(preds, y):
        tgt = y.squeeze().float()
        return sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0]

    with cpu_rand():
        # ---- warm-up
        domain_list = list(range(n_domain))
        for _ in range(cfg["warm_up"]):
            if not domain_list:
                domain_list = list(range(n_domain))
            d = domain_list.pop()
            x, y = to_dev(*get(d))
            pred = model(x, mode="wo_mask", domain_i=d, memory_gate_value=True)
            loss = crit(pred.squeeze(), y.squeeze().float()) + model.get_regularization_loss(device=device)
            model.zero_grad()
            loss.backward()
            opt.step()
            out["warm_up"].append(float(loss))

        probe = ("tower_gates.", "mmoe_gates.", "group_embedding.", "linear.", "cn.", "towers_linear.", "towers.2.0.")
        out["after_warm_up"] = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()
                                if k.startswith(probe) and v.numel() <= 4096 and v.is_floating_point()}

        # ---- regroup
        model.save_model_state()
        sigma = cfg["random_modify_sigma"] * 0.99
        active = max(0.1, cfg["init_active_percent"] * 0.95)
        for d in range(n_domain):
            for z in range(cfg["candidates"]):
                tmp_mask = model.generate_mask(generate_mode="mask_max_gate", d=d, init_active_percent=active,
                                               random_modify_sigma=sigma)
                generated = mask_to_lists(tmp_mask)
                model.load_model_state()
                fast = torch.optim.Adam(model.parameters(), lr=UPDATE_LR, betas=(0.9, 0.99), eps=1e-8, weight_decay=WD)
                upd = []
                for _ in range(cfg["update_steps"]):
                    x, y = to_dev(*get(d, "aug_train"))
                    preds = model(x, mode="domain_mask_bagging", current_mask=tmp_mask, tmp_memory_gate_value=True)
                    loss = bagging(preds, y) + model.get_regularization_loss(device=device)
                    model.zero_grad()
                    loss.backward()
                    fast.step()
                    upd.append(float(loss))
                    out["gate_log"].append([torch.stack(model.tmp_tower_gate_values[l], dim=1).detach().cpu().clone()
                                            for l in range(1, spec.n_level)])
                    tmp_mask = model.prun_single_mask(d, tmp_mask, prun_ratio=0.05)
                    out["prune_log"].append(mask_to_lists(tmp_mask))
                model.candidate_domain_mask[d].append(tmp_mask)
                out["update"].append(upd)
                out["candidates"].append({"d": d, "z": z, "generated": generated, "pruned": mask_to_lists(tmp_mask)})
                ev = []
                with torch.no_grad():
                    for _ in range(cfg["eval_steps"]):
                        x, y = to_dev(*get(d))
                        pred = model(x, mode="domain_with_mask", current_mask=tmp_mask)
                        loss = crit(pred.squeeze(), y.squeeze().float()) + model.get_regularization_loss(device=device)
                        model.add_eval_loss(loss.mean().item(), d=d, mask_z=z)
                        ev.append(float(loss))
                out["eval_loss"].append(ev)
        model.update_all_mask(regroup_times=1)
        model.reset_for_mask_update()
        model.load_model_state()
        out["selected"] = [mask_to_lists(model.domain_mask[d]) for d in range(n_domain)]

        # ---- training under the selected masks
        for i in range(cfg["post_steps"]):
            d = i % n_domain
            x, y = to_dev(*get(d))
            preds = model(x, mode="domain_mask_bagging", domain_i=d, memory_gate_value=(i % 2 == 1))
            loss = bagging(preds, y) + model.get_regularization_loss(device=device)
            model.zero_grad()
            loss.backward()
            opt.step()
            out["post"].append(float(loss))
        out["recorded_gate_counts"] = [[[len(model.domain_tower_gate_values[d][l][t]) for t in range(spec.n_tower[l])]
                                        for l in range(spec.n_level)] for d in range(n_domain)]
    return out
