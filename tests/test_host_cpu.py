"""Host pieces that need no GPU: packed parameter storage (packing.py) on a CPU-built model, the synthetic workloads
of bench.py, and the loud failure of the compute path off-GPU (there is no CPU fallback)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import aread_torch as O
from tests._models import build_model
from tests._util import load_golden

workloads = importlib.import_module("aread-multi-domain-recommendation_b200.workloads")


def _model():
    fx = load_golden("tiny")
    return fx, build_model(O.Spec(**fx["spec"]), "cpu")


def test_packed_storage_backs_the_parameters():
    fx, model = _model()
    names = dict(model.named_parameters())
    for pack in model._packs.packs:
        flat = pack.flat
        members = pack.fetch()
        assert flat.shape[0] == len(members)
        for i, p in enumerate(members):
            assert p.data_ptr() == flat[i].data_ptr() and p.shape == flat[i].shape    # a view, not a copy
    # state_dict keeps the reference's per-module names and shapes; writing through it reaches the packs
    sd = model.state_dict()
    key = "mmoe_experts.1.layers.0.weight"
    assert key in sd and key in names
    with torch.no_grad():
        names[key].fill_(0.25)
    assert float(model._expert_layers[0].weight.flat[1].mean()) == 0.25
    new = {k: torch.full_like(v, 0.5) if k == key else v for k, v in sd.items()}
    model.load_state_dict(new)
    assert float(model._expert_layers[0].weight.flat[1].mean()) == 0.5
    # deepcopy re-packs its own storage
    import copy
    clone = copy.deepcopy(model)
    assert clone._packs is not model._packs
    assert clone._expert_layers[0].weight.flat.data_ptr() != model._expert_layers[0].weight.flat.data_ptr()
    assert torch.equal(clone.state_dict()[key], model.state_dict()[key])


def test_compute_path_refuses_the_cpu():
    fx, model = _model()
    x = torch.zeros((4, len(model.embedding.offsets)), dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        model(x, mode="wo_mask")
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        model.get_regularization_loss(device=torch.device("cpu"))


@pytest.mark.parametrize("name", ["amazon", "aliccp", "cloudtheme"])
def test_workloads_are_deterministic_and_in_range(name):
    wl = workloads.WORKLOADS[name]()
    x, y, d = wl.batch(512, seed=7)
    x2, y2, d2 = wl.batch(512, seed=7)
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and d == d2
    assert x.dtype == np.int32 and x.shape == (512, wl.n_cols) and y.shape == (512, 1)
    dims = np.asarray(wl.one_hot_field_dims)
    n_one_hot = len(dims)
    assert (x[:, :n_one_hot] >= 0).all() and (x[:, :n_one_hot] < dims[None, :]).all()
    assert (x[:, wl.domain_idx] == d).all() and 0 <= d < wl.n_domain          # single-domain batch, like the loaders
    if wl.n_cols > n_one_hot:                                                  # history columns: item ids or the pad id
        hist = x[:, n_one_hot:]
        assert (hist >= 0).all() and (hist <= dims[wl.itemid_idx]).all()
    assert wl.gather_bytes_per_sample() > 0 and wl.n_rows == int(dims.sum())
    assert len({wl.batch(8, seed=s)[2] for s in range(40)}) > 1               # domains vary from batch to batch
