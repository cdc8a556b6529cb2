"""Host-side HEMP (Hierarchical Expert Mask Pruning) bookkeeping.

Pure numpy: masks are tiny (105 edges at the default sizes) and this logic runs once per regroup,
never inside a step.  Behaviour -- including the order in which `np.random` / `torch.rand` are
consumed -- follows the reference `model/aread.py`:
  create_single_full_mask :548-568   validate_mask :570-605   count_active_edge :671-680
A mask is a list of n_level + 1 boolean matrices: mask[0] is [1, n_tower[0]] (bottom inputs),
mask[l] is [n_tower[l-1], n_tower[l]] (edges into level l) and mask[-1] is [n_tower[-1], 1].
"""
import numpy as np
import torch


def level_shapes(n_tower):
    n_tower = tuple(n_tower)
    return [(1, n_tower[0])] + [(n_tower[l - 1], n_tower[l]) for l in range(1, len(n_tower))] + [(n_tower[-1], 1)]


def full_mask(n_tower, fill_value=0):
    """All-False, all-True or Bernoulli(fill_value) mask as numpy arrays.  The random variant draws
    with np.random.choice level by level, like the reference, so a seeded run yields the same mask."""
    shapes = level_shapes(n_tower)
    if fill_value == 0:
        return [np.zeros(s, dtype=bool) for s in shapes]
    if fill_value == 1:
        return [np.ones(s, dtype=bool) for s in shapes]
    if 0 < fill_value < 1:
        return [np.random.choice([True, False], s, p=[fill_value, 1 - fill_value]) for s in shapes]
    raise ValueError('fill_value in mask must be 0 or 1 or (0, 1)')


def to_numpy(mask):
    return [m.detach().cpu().numpy().copy() if isinstance(m, torch.Tensor) else np.array(m, dtype=bool) for m in mask]


def validate_arrays(m, n_tower, add_input=True, add_output=True, remove_hidden=True):
    """In-place repair of a numpy mask: bottom towers that feed something get their input edge, top
    towers that are fed get their output edge, and towers without inputs (outputs) lose their
    outputs (inputs), propagating downwards through a work queue."""
    n_level = len(n_tower)
    if add_input:
        m[0][0, m[1].any(axis=1)] = True
    if add_output:
        m[-1][m[-2].any(axis=0), 0] = True
    if remove_hidden:
        queue = [(l, t) for l in range(1, n_level) for t in range(n_tower[l])]
        while queue:
            l, t = queue.pop(0)
            if not m[l][:, t].any():
                m[l + 1][t, :] = False
            if not m[l + 1][t, :].any():
                col = t
                if l > 1:
                    feeders = np.nonzero(m[l][:, t])[0].tolist()
                    for src in feeders:
                        if (l - 1, src) not in queue:
                            queue.append((l - 1, src))
                    if feeders:
                        # the reference re-uses its loop variable here (aread.py:601-604), so the column
                        # that is cleared is the last feeder's index, not the tower's own; kept as is
                        col = feeders[-1]
                m[l][:, col] = False
    return m


def validate(mask, n_tower, add_input=True, add_output=True, remove_hidden=True):
    """validate_mask with the reference's in-place contract: the same list (and the same tensors or
    arrays inside it) comes back modified."""
    arrays = to_numpy(mask)
    validate_arrays(arrays, tuple(n_tower), add_input, add_output, remove_hidden)
    for i, a in enumerate(arrays):
        if isinstance(mask[i], torch.Tensor):
            mask[i].copy_(torch.from_numpy(a))
        else:
            mask[i][...] = a
    return mask


def count_edges(mask):
    total = 0
    for m in mask:
        total += torch.sum(m).cpu().item() if isinstance(m, torch.Tensor) else np.sum(m)
    return total


def as_device_mask(arrays, device):
    return [torch.tensor(a, dtype=torch.bool, device=device) for a in arrays]
