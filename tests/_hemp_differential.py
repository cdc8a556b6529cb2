"""Run in a subprocess by tests/test_hemp_reference_cpu.py: the UNMODIFIED reference AREAD (imported from
$AREAD_REF, first on sys.path) against this repository's HEMP host logic on many random masks and seeds."""
import importlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AREAD_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REF, ROOT]                      # `model.aread` -> the reference; the package is imported by its own name

from tests.golden import make_golden as G      # noqa: E402

hemp = importlib.import_module("aread-multi-domain-recommendation_b200.hemp")
mine_mod = importlib.import_module("aread-multi-domain-recommendation_b200.aread")


def build_mine(spec):
    mh = {"multi_hot_flag": list(spec.flag), "itemid_idx": spec.itemid_idx, "seq_maxlen": spec.seq_maxlen,
          "method": spec.method}
    from tests._models import make_config
    m = mine_mod.AREAD(np.asarray(spec.one_hot_field_dims), spec.embed_dim, mh, n_tower=tuple(spec.n_tower),
                       n_domain=spec.n_domain, base_model="mmoe", expert_dims=tuple(spec.expert_dims),
                       tower_dims=tuple(tuple(t) for t in spec.tower_dims), domain_idx=spec.domain_idx,
                       device=torch.device("cpu"), dropout=0.0, config=make_config(spec))
    m.reset_for_mask_update()
    return m


def main():
    refcfg, RefAREAD = G.load_reference()
    assert RefAREAD.__module__ == "model.aread" and REF in sys.modules["model.aread"].__file__
    checked = 0
    for case in ("ali_small", "tiny"):
        spec = G.Spec(**G.CASES[case]["spec"])
        ref = G.build_reference(refcfg, RefAREAD, spec, dropout=0.0).eval()
        mine = build_mine(spec)
        nt = spec.n_tower
        rng = np.random.RandomState(1234)
        for i in range(300):                     # validate_mask: numpy and tensor inputs, every flag combination
            p = rng.choice([0.05, 0.15, 0.3, 0.5, 0.8])
            raw = [rng.rand(1, nt[0]) < p] + [rng.rand(nt[l - 1], nt[l]) < p for l in range(1, len(nt))] + \
                  [rng.rand(nt[-1], 1) < p]
            flags = dict(add_input=bool(i & 1), add_output=bool(i & 2), remove_hidden=bool(i & 4) or i % 3 == 0)
            if i % 2:
                a = ref.validate_mask([torch.tensor(r) for r in raw], **flags)
                b = mine.validate_mask([torch.tensor(r) for r in raw], **flags)
            else:
                a = ref.validate_mask([r.copy() for r in raw], **flags)
                b = mine.validate_mask([r.copy() for r in raw], **flags)
            for l, (u, v) in enumerate(zip(a, b)):
                assert np.array_equal(np.asarray(u), np.asarray(v)), (case, i, l, flags)
            checked += 1
        for seed in range(40):                   # generate_mask('rand'): same np.random consumption, same masks
            p = [0.9, 0.7, 0.5, 0.3, 0.15, 0.1][seed % 6]
            np.random.seed(seed)
            a = ref.generate_mask("rand", seed % spec.n_domain, init_active_percent=p)
            after_a = np.random.rand()
            np.random.seed(seed)
            b = mine.generate_mask("rand", seed % spec.n_domain, init_active_percent=p)
            after_b = np.random.rand()
            assert after_a == after_b, (case, seed)
            for l, (u, v) in enumerate(zip(a, b)):
                assert torch.equal(u.cpu(), v.cpu()), (case, seed, l)
            assert ref.count_active_edge(0, a) == mine.count_active_edge(0, b)
            checked += 1
    print("HEMP DIFFERENTIAL OK", checked)


if __name__ == "__main__":
    main()
