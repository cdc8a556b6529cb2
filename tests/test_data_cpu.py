"""data.py: the device batch loader visits rows in exactly the order of the trainer's
`DataLoader(TensorDataset(...), bs, shuffle=True)` under the same seed, and the .pth caches round-trip."""
import importlib

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

data = importlib.import_module("aread-multi-domain-recommendation_b200.data")


@pytest.mark.parametrize("n,bs", [(1000, 128), (37, 64), (64, 64), (5, 1)])
def test_same_batches_as_dataloader(n, bs):
    X = torch.arange(n * 3, dtype=torch.int32).reshape(n, 3)
    y = (torch.arange(n) % 2).to(torch.int16).reshape(n, 1)
    torch.manual_seed(2000)
    ref = []
    ref_loader = DataLoader(TensorDataset(X, y), bs, shuffle=True)
    for _ in range(2):                                      # two epochs: the RNG keeps advancing the same way
        ref.append([[t.clone() for t in batch] for batch in ref_loader])
    after_ref = torch.rand(1)
    torch.manual_seed(2000)
    mine_loader = data.DeviceBatchLoader(X, y, batch_size=bs, shuffle=True)
    assert len(mine_loader) == len(ref_loader) and len(mine_loader.dataset) == len(ref_loader)
    for epoch in range(2):
        got = list(mine_loader)
        assert len(got) == len(ref[epoch])
        for a, b in zip(got, ref[epoch]):
            assert len(a) == len(b) == 2
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
            assert a[0].dtype == torch.int32 and a[1].dtype == torch.int16
    assert torch.equal(torch.rand(1), after_ref)            # consumed the global RNG identically


def test_unshuffled_and_domain_loaders():
    n = 50
    X = torch.stack([torch.arange(n), torch.arange(n) % 4], dim=1).to(torch.int32)
    y = torch.zeros(n, 1, dtype=torch.int16)
    batches = list(data.DeviceBatchLoader(X, y, batch_size=16, shuffle=False))
    assert [b[0].shape[0] for b in batches] == [16, 16, 16, 2]
    assert torch.equal(torch.cat([b[0] for b in batches]), X)
    loaders, seq = data.domain_loaders(X, y, 1, [0, 2, 3], 8)
    assert seq == [0, 0, 2, 2, 3, 3]                        # 13, 12, 12 rows -> 2 batches each
    for d, ld in zip([0, 2, 3], loaders):
        rows = torch.cat([b[0] for b in ld])
        assert (rows[:, 1] == d).all() and rows.shape[0] == int((X[:, 1] == d).sum())
        assert sorted(rows[:, 0].tolist()) == X[X[:, 1] == d][:, 0].tolist()
    with pytest.raises(ValueError):
        data.DeviceBatchLoader(X, y[:-1], batch_size=4)


def test_split_files_round_trip(tmp_path):
    X = torch.randint(0, 1 << 20, (123, 17), dtype=torch.int32)
    y = torch.randint(0, 2, (123, 1)).to(torch.int16)
    data.save_split(str(tmp_path), "train", X, y)
    # what run.py:262-263 writes is readable, and what we write is what run.py:274-275 reads
    assert torch.equal(torch.load(tmp_path / "train_data_loader.pth"), X)
    assert torch.equal(torch.load(tmp_path / "train_label_loader.pth"), y)
    X2, y2 = data.load_split(str(tmp_path), "train")
    assert torch.equal(X2, X) and torch.equal(y2, y) and X2.dtype == torch.int32 and y2.dtype == torch.int16
    torch.save(X.to(torch.int64), tmp_path / "valid_data_loader.pth")      # older caches: converted on load
    torch.save(y.to(torch.int64), tmp_path / "valid_label_loader.pth")
    X3, y3 = data.load_split(str(tmp_path), "valid")
    assert torch.equal(X3, X) and torch.equal(y3, y) and X3.dtype == torch.int32
    with pytest.raises(TypeError):
        data.save_split(str(tmp_path), "test", X.to(torch.int64), y)


def test_eval_accumulator_matches_per_batch_host_copies():
    import numpy as np
    acc = data.EvalAccumulator()
    fixed = data.EvalAccumulator(capacity=100)
    want = [[], [], []]
    g = torch.Generator().manual_seed(0)
    for n in (7, 1, 32):
        y = (torch.rand(n, 1, generator=g) < 0.3).to(torch.int16)
        p = torch.rand(n, generator=g)
        d = torch.randint(0, 5, (n,), generator=g, dtype=torch.int32)
        acc.add(y, p, d)
        fixed.add(y, p, d)
        want[0].append(y.squeeze(-1).numpy())       # what run.py:725-727 collects batch by batch
        want[1].append(p.numpy())
        want[2].append(d.numpy())
    for got in (acc.result(), fixed.result()):
        for a, b in zip(got, want):
            assert np.array_equal(a, np.concatenate(b).astype(a.dtype))
    with pytest.raises(ValueError):
        fixed.add(torch.zeros(100), torch.zeros(100), torch.zeros(100, dtype=torch.int32))
