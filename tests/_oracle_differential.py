"""Run in a subprocess by tests/test_oracle_reference_cpu.py: the oracle (oracle/aread_torch.py) against the
UNMODIFIED reference AREAD on RANDOM model shapes, batches and masks -- beyond the three committed fixtures."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AREAD_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REF, ROOT]

from oracle import aread_torch as O            # noqa: E402
from oracle import synth                       # noqa: E402
from tests.golden import make_golden as G      # noqa: E402

RTOL, ATOL = 5e-5, 5e-6


def close(a, b, what, rtol=RTOL, atol=ATOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), (what, float(err.max()))


def random_spec(rng):
    n_level = 3
    n_tower = tuple(int(v) for v in rng.randint(2, 6, size=n_level))
    widths = [int(rng.choice([4, 8, 12])) for _ in range(n_level)]
    h = int(rng.choice([8, 16]))
    tower_dims = []
    prev = h
    for l in range(n_level):
        tower_dims.append((int(rng.choice([8, 12, 16])), widths[l]))
        prev = widths[l]
    n_one_hot = int(rng.randint(3, 7))
    dims = [int(v) for v in rng.randint(3, 60, size=n_one_hot)]
    domain_idx = int(rng.randint(0, n_one_hot))
    kw = dict(one_hot_field_dims=dims, embed_dim=int(rng.choice([4, 8])), n_domain=dims[domain_idx],
              domain_idx=domain_idx, itemid_idx=0, n_tower=n_tower, expert_dims=(int(rng.choice([16, 24])), h),
              tower_dims=tuple(tower_dims), n_expert=int(rng.randint(2, 5)), n_cross_layers=int(rng.randint(1, 4)))
    if rng.rand() < 0.5:                           # history fields pooled over seq_maxlen item ids
        n_mh, L = int(rng.randint(1, 3)), int(rng.randint(2, 5))
        kw.update(multi_hot_flag=[False] * n_one_hot + [True] * (n_mh * L), seq_maxlen=L,
                  method=str(rng.choice(["mean", "sum"])))
    return O.Spec(**kw)


def main():
    refcfg, RefAREAD = G.load_reference()
    assert REF in sys.modules["model.aread"].__file__
    rng = np.random.RandomState(2026)
    n = 0
    for case in range(6):
        spec = random_spec(rng)
        B = int(rng.choice([1, 2, 9, 33]))
        dom = int(rng.randint(0, spec.n_domain))
        pad = spec.one_hot_field_dims[spec.itemid_idx] if spec.n_mh_fields and rng.rand() < 0.5 else None
        x, y = synth.random_batch(spec, B, seed=100 + case, domain=dom, pad_id=pad)
        ref = G.build_reference(refcfg, RefAREAD, spec, dropout=0.0)
        np.random.seed(case)
        mask = ref.generate_mask("rand", 0, init_active_percent=float(rng.choice([0.3, 0.6, 0.9])))
        sd = synth.deterministic_state(spec)
        # eval forwards
        ref.eval()
        with torch.no_grad():
            close(O.embed(sd, spec, x), ref.embedding(x), f"{case} embed", 0, 0)
            close(O.forward(sd, spec, x, "wo_mask")["y"], ref(x, mode="wo_mask"), f"{case} wo_mask")
            close(O.forward(sd, spec, x, "domain_with_mask", mask)["y"],
                  ref(x, mode="domain_with_mask", current_mask=[t.clone() for t in mask]), f"{case} with_mask")
            close(O.reg_loss(sd, spec), ref.get_regularization_loss(device=torch.device("cpu")), f"{case} reg", 1e-6, 0)
        # one train step: bagging loss + L2, gradients of every parameter that has one
        ref.train()
        preds = ref(x, mode="domain_mask_bagging", current_mask=[t.clone() for t in mask])
        tgt = y.reshape(-1).float()
        crit = torch.nn.BCELoss()
        loss = sum(crit(p, tgt) for p in preds.unbind(0)) / preds.shape[0] + \
            ref.get_regularization_loss(device=torch.device("cpu"))
        ref.zero_grad()
        loss.backward()
        leaves = O.make_leaf_params(synth.deterministic_state(spec))
        out = O.forward(leaves, spec, x, "domain_mask_bagging", mask, training=True)
        mine = O.bagging_loss(out["y"], y) + O.reg_loss(leaves, spec)
        mine.backward()
        close(out["y"], preds.detach(), f"{case} y_stack")
        close(mine.detach(), loss.detach(), f"{case} loss")
        dead = ("atten_", "self_attns", "V_res", "final_gate")
        for k, p in ref.named_parameters():
            if k.startswith(dead):
                continue
            g = leaves[k].grad
            assert (g is None) == (p.grad is None), (case, k)
            if g is not None:
                scale = float(p.grad.abs().max()) + 1e-12
                assert float((g - p.grad).abs().max()) <= 2e-4 * scale + 2e-7, (case, k)
        n += 1
    print("ORACLE DIFFERENTIAL OK", n)


if __name__ == "__main__":
    main()
