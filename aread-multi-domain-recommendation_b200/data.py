"""Batch feeding and on-disk tensors of the trainer's data path.

`DeviceBatchLoader` is a drop-in for `DataLoader(TensorDataset(X, y[, group]), bs, shuffle=True)` over tensors
that already live on the device (run.py:301-306, 334-335).  The stock loader collates a batch from `bs`
single-row index operations (1024 tiny kernels per tensor per batch); here a batch is one `index_select`
per tensor with a slice of a device-resident permutation.  The permutation is drawn exactly as
`RandomSampler` draws it (same global-RNG consumption, same order), so a run is reproducible against the
reference batch for batch -- pinned in tests/test_data_cpu.py.

`save_split` / `load_split`: the `{mode}_data_loader.pth` (int32 [N, n_cols]) and `{mode}_label_loader.pth`
(int16 [N, 1]) caches under `dataset/<name>/<csv-stem>/` (run.py:262-263, 274-279), byte-compatible with
`torch.save` of the tensors."""
import math
import os

import torch


class DeviceBatchLoader:
    def __init__(self, *tensors, batch_size=1, shuffle=True, generator=None):
        if not tensors:
            raise ValueError("at least one tensor is required")
        n = tensors[0].shape[0]
        if any(t.shape[0] != n for t in tensors):
            raise ValueError("Size mismatch between tensors")               # TensorDataset's message
        self.tensors = tensors
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.generator = generator
        self.dataset = self                                                  # `len(loader.dataset)` keeps working

    def __len__(self):
        return math.ceil(self.tensors[0].shape[0] / self.batch_size)

    @property
    def n_rows(self):
        return self.tensors[0].shape[0]

    def _permutation(self):
        """The order torch's DataLoader(shuffle=True, num_workers=0) would visit the rows in, consuming the
        global RNG the same way: one draw for the iterator's base seed, one for the sampler's generator."""
        n = self.n_rows
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)        # _BaseDataLoaderIter base seed
        if self.generator is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
        else:
            g = self.generator
        return torch.randperm(n, generator=g)

    def __iter__(self):
        n, bs = self.n_rows, self.batch_size
        dev = self.tensors[0].device
        if self.shuffle:
            perm = self._permutation().to(dev, non_blocking=True)
            for s in range(0, n, bs):
                idx = perm[s:s + bs]
                yield [t.index_select(0, idx) for t in self.tensors]
        else:
            for s in range(0, n, bs):
                yield [t[s:s + bs] for t in self.tensors]


def domain_loaders(X, y, domain_idx, domains, batch_size, shuffle=True):
    """Per-domain loaders and the batch sequence of run.py:320-335: ([loader per domain], [d repeated
    ceil(rows_d / bs) times])."""
    loaders, seq = [], []
    col = X[:, domain_idx]
    for d in domains:
        rows = (col == d).nonzero(as_tuple=True)[0]
        loaders.append(DeviceBatchLoader(X.index_select(0, rows), y.index_select(0, rows), batch_size=batch_size,
                                         shuffle=shuffle))
        seq.extend([d] * math.ceil(rows.numel() / batch_size))
    return loaders, seq


def save_split(folder, mode, X, y):
    if X.dtype != torch.int32 or y.dtype != torch.int16:
        raise TypeError(f"ids must be int32 and labels int16 (run.py:255-259), got {X.dtype} / {y.dtype}")
    if X.dim() != 2 or y.shape[0] != X.shape[0]:
        raise ValueError(f"expected ids [N, n_cols] and N labels, got {tuple(X.shape)} / {tuple(y.shape)}")
    os.makedirs(folder, exist_ok=True)
    torch.save(X.cpu(), os.path.join(folder, f"{mode}_data_loader.pth"))
    torch.save(y.cpu(), os.path.join(folder, f"{mode}_label_loader.pth"))


def load_split(folder, mode, device=None):
    """(ids int32 [N, n_cols], labels int16 [N, 1]) as run.py:274-275 reads them."""
    X = torch.load(os.path.join(folder, f"{mode}_data_loader.pth")).to(torch.int32)
    y = torch.load(os.path.join(folder, f"{mode}_label_loader.pth")).to(torch.int16)
    if X.dim() != 2 or y.shape[0] != X.shape[0]:
        raise ValueError(f"{mode}: ids {tuple(X.shape)} and labels {tuple(y.shape)} do not match")
    if device is not None:
        X, y = X.to(device), y.to(device)
    return X, y


class EvalAccumulator:
    """Targets / predictions / domain ids of an evaluation pass kept ON THE DEVICE and brought to the host once.

    The reference's test loop (run.py:725-727, 742-744) calls `.cpu().numpy()` three times per batch -- three stream
    synchronisations that serialise every batch with the host.  Here `add` only keeps references (or appends into
    preallocated device buffers when `capacity` is given) and `result()` does one concatenation and one copy; the
    numpy arrays it returns are what `roc_auc_score` / `log_loss` are fed at run.py:757-758."""

    def __init__(self, capacity=None, device=None):
        self._chunks = []
        self._buf = None
        self._n = 0
        if capacity is not None:
            self._buf = (torch.empty(capacity, dtype=torch.float32, device=device),
                         torch.empty(capacity, dtype=torch.float32, device=device),
                         torch.empty(capacity, dtype=torch.int32, device=device))

    def add(self, targets, predicts, domains):
        t, p, d = targets.reshape(-1), predicts.reshape(-1), domains.reshape(-1)
        if self._buf is None:
            self._chunks.append((t, p, d))
        else:
            n = t.numel()
            if self._n + n > self._buf[0].numel():
                raise ValueError("EvalAccumulator capacity exceeded")
            self._buf[0][self._n:self._n + n].copy_(t)
            self._buf[1][self._n:self._n + n].copy_(p)
            self._buf[2][self._n:self._n + n].copy_(d)
        self._n += t.numel()

    def result(self):
        """(targets, predicts, domains) as numpy arrays, in the order they were added."""
        if self._buf is not None:
            parts = [b[:self._n] for b in self._buf]
        else:
            parts = [torch.cat([c[i].to(torch.float32 if i < 2 else torch.int32) for c in self._chunks])
                     if self._chunks else torch.empty(0) for i in range(3)]
        return tuple(p.cpu().numpy() for p in parts)
