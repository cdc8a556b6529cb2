"""Contiguous storage for per-expert / per-tower parameters.

The reference keeps one nn.Linear / nn.BatchNorm1d per expert and per tower (state_dict keys
`mmoe_experts.{e}.layers.{i}.*`, `towers.{l}.{t}.layers.{i}.*`, ...).  The grouped kernels want the
groups of a layer side by side, so every such family is re-pointed at slices of one packed tensor:
the Parameter objects, their names and shapes stay what the reference has (optimizers,
`state_dict`, `load_state_dict`, checkpoints keep working) while `pack.flat` is the [groups, ...]
array the kernels read.  `nn.Module._apply` (`.to()`, `.cuda()`, `.float()`) replaces `.data`, so the
model re-packs after it.
"""
import torch


class Pack:
    def __init__(self, fetch):
        """`fetch()` returns the current tensor objects (nn.Module._apply replaces buffer objects, so
        they are looked up again at every re-pack)."""
        self.fetch = fetch if callable(fetch) else (lambda tensors=list(fetch): tensors)
        self.tensors = list(self.fetch())
        shapes = {tuple(t.shape) for t in self.tensors}
        if len(shapes) != 1:
            raise ValueError(f"cannot pack tensors of different shapes: {shapes}")
        self.shape = tuple(self.tensors[0].shape)
        self.flat = None
        self.repack()

    def repack(self):
        self.tensors = list(self.fetch())
        t0 = self.tensors[0]
        flat = torch.empty((len(self.tensors),) + self.shape, dtype=t0.dtype, device=t0.device)
        with torch.no_grad():
            for i, t in enumerate(self.tensors):
                flat[i].copy_(t.data)
                t.data = flat[i]
        self.flat = flat

    def intact(self):
        self.tensors = list(self.fetch())
        f = self.flat
        if f is None or f.device != self.tensors[0].device:
            return False
        step = f.stride(0) * f.element_size() if f.dim() > 0 and f.shape[0] > 1 else 0
        base = f.data_ptr()
        return all(t.data_ptr() == base + i * step for i, t in enumerate(self.tensors))

    @property
    def groups(self):
        return len(self.tensors)

    def rows(self):
        """[groups * shape[0], *shape[1:]] view: the groups stacked by rows."""
        return self.flat.view((self.groups * self.shape[0],) + self.shape[1:]) if self.shape else self.flat


class PackSet:
    """All packs of one model; `ensure()` re-packs whatever `.to()` / deepcopy un-aliased."""

    def __init__(self):
        self.packs = []

    def add(self, fetch):
        pack = Pack(fetch)
        self.packs.append(pack)
        return pack

    def ensure(self):
        for pack in self.packs:
            if not pack.intact():
                pack.repack()
