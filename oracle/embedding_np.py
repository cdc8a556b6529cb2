"""numpy restatement of the multi-field embedding lookup (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows `/root/reference/model/layer.py`:
  * FeaturesEmbedding.__init__  layer.py:131-157  -> `field_offsets`
  * FeaturesEmbedding.forward   layer.py:160-183  -> `lookup_rows`, `gather_fwd`
  * autograd of the above (aten::embedding_dense_backward, dense because
    `sparse=False`, layer.py:150)              -> `scatter_bwd_dense`
and states the index bookkeeping the CUDA scatter uses (`sort_segments`) so that it can
be checked bit for bit.
"""
import numpy as np


def field_offsets(one_hot_field_dims, multi_hot_flag=None, itemid_idx=0):
    """Row offset of every input column (layer.py:151-157).

    One-hot columns get the exclusive cumsum of the field dims; every multi-hot column
    re-uses the offset of the item-id field.  Offsets are positional: column c of `x`
    is shifted by offsets[c].
    """
    dims = np.asarray(one_hot_field_dims, dtype=np.int64)
    off = np.zeros(len(dims), dtype=np.int64)
    if len(dims) > 1:
        off[1:] = np.cumsum(dims)[:-1]
    n_mh = int(np.sum(np.asarray(multi_hot_flag, dtype=bool))) if multi_hot_flag is not None else 0
    if n_mh > 0:
        off = np.concatenate([off, np.full(n_mh, off[itemid_idx], dtype=np.int64)])
    return off


def lookup_rows(x, offsets, n_rows):
    """idx = x + offsets in x's own integer type (layer.py:165), then the bounds rule of
    F.embedding: only idx outside [0, n_rows) raises; an id beyond its own field silently
    aliases into the next field's rows."""
    x = np.asarray(x)
    idx = (x + np.asarray(offsets).astype(x.dtype)[None, :]).astype(x.dtype)
    if idx.size and (idx.min() < 0 or idx.max() >= n_rows):
        raise IndexError("index out of range in self")
    return idx


def gather_fwd(weight, x, offsets, multi_hot_flag=None, seq_maxlen=1, method=None):
    """[B, n_cols] ids -> [B, output_dim0, D] fp32 (layer.py:165-178).

    Pooled multi-hot fields are summed over the `seq_maxlen` positions in order
    (padding positions included) and, for 'mean', divided by seq_maxlen afterwards.
    """
    weight = np.asarray(weight, dtype=np.float32)
    idx = lookup_rows(x, offsets, weight.shape[0])
    e = weight[idx.astype(np.int64)]                      # [B, n_cols, D]
    flag = None if multi_hot_flag is None else np.asarray(multi_hot_flag, dtype=bool)
    if flag is None or not flag.any() or method not in ("mean", "sum"):
        return e
    B, _, D = e.shape
    one_hot = e[:, ~flag, :]
    mh = e[:, flag, :].reshape(B, int(flag.sum()) // seq_maxlen, seq_maxlen, D)
    acc = mh[:, :, 0, :].copy()
    for l in range(1, seq_maxlen):
        acc = acc + mh[:, :, l, :]
    if method == "mean":
        acc = acc / np.float32(seq_maxlen)
    return np.concatenate([one_hot, acc.astype(np.float32)], axis=1)


def expand_pooled_grad(d_out, n_cols, multi_hot_flag=None, seq_maxlen=1, method=None, reciprocal=False):
    """Gradient of the pooling/concat: [B, output_dim0, D] -> per-column [B, n_cols, D].  The
    reference divides by seq_maxlen; the CUDA scatter multiplies by fl32(1/seq_maxlen)
    (`reciprocal=True`), which can differ in the last bit."""
    d_out = np.asarray(d_out, dtype=np.float32)
    flag = None if multi_hot_flag is None else np.asarray(multi_hot_flag, dtype=bool)
    if flag is None or not flag.any() or method not in ("mean", "sum"):
        return d_out
    B, _, D = d_out.shape
    n_oh = int((~flag).sum())
    g = np.empty((B, n_cols, D), dtype=np.float32)
    g[:, ~flag, :] = d_out[:, :n_oh, :]
    gp = d_out[:, n_oh:, :]
    if method == "mean":
        gp = gp * (np.float32(1.0) / np.float32(seq_maxlen)) if reciprocal else gp / np.float32(seq_maxlen)
    g[:, flag, :] = np.repeat(gp, seq_maxlen, axis=1)
    return g


def scatter_bwd_dense(d_out, x, offsets, n_rows, multi_hot_flag=None, seq_maxlen=1, method=None):
    """Dense [R, D] table gradient; duplicates are accumulated in flattened (b, c) order,
    which is the order aten::embedding_dense_backward uses on the CPU."""
    idx = lookup_rows(x, offsets, n_rows).astype(np.int64)
    g = expand_pooled_grad(d_out, idx.shape[1], multi_hot_flag, seq_maxlen, method)
    dw = np.zeros((n_rows, g.shape[-1]), dtype=np.float32)
    np.add.at(dw, idx.reshape(-1), g.reshape(-1, g.shape[-1]))
    return dw


def sort_segments(idx_flat):
    """Bookkeeping of the sort-by-row segmented reduce.

    Returns (sorted_rows, perm, unique_rows, seg_offsets): a *stable* sort of the lookups by
    table row, the position each sorted entry came from, the distinct rows in ascending
    order and the start of every row's segment (len = n_unique + 1)."""
    idx_flat = np.asarray(idx_flat).reshape(-1).astype(np.int64)
    perm = np.argsort(idx_flat, kind="stable")
    sorted_rows = idx_flat[perm]
    if len(sorted_rows) == 0:
        return sorted_rows, perm, sorted_rows, np.zeros(1, dtype=np.int64)
    head = np.ones(len(sorted_rows), dtype=bool)
    head[1:] = sorted_rows[1:] != sorted_rows[:-1]
    starts = np.nonzero(head)[0]
    return sorted_rows, perm, sorted_rows[starts], np.concatenate([starts, [len(sorted_rows)]]).astype(np.int64)


def scatter_bwd_tiled(d_cols, idx_flat, n_rows, tile=32):
    """The summation order of the CUDA segmented reduce (include/aread_sm100.h): the stably sorted
    lookups are covered by an aligned `tile`-ary tree whose level-l blocks hold tile**l consecutive
    entries.  A row's entries inside one level-1 block are summed left to right; its sum inside a
    level-l block is the left to right sum of its sums inside the block's children.  A row whose
    entries all fall inside one level-1 block reproduces the reference (sequential) order."""
    idx_flat = np.asarray(idx_flat).reshape(-1)
    d_cols = np.asarray(d_cols, dtype=np.float32).reshape(len(idx_flat), -1)
    _, perm, uniq, seg = sort_segments(idx_flat)
    top = 1
    while tile ** top < max(len(idx_flat), 1):
        top += 1

    def part(s, e, level):
        if level == 1:
            acc = d_cols[perm[s]].copy()
            for p in perm[s + 1:e]:
                acc = acc + d_cols[p]
            return acc
        child = tile ** (level - 1)
        acc = None
        c0 = s
        while c0 < e:
            c1 = min(e, (c0 // child + 1) * child)
            sub = part(c0, c1, level - 1)
            acc = sub if acc is None else acc + sub
            c0 = c1
        return acc

    dw = np.zeros((n_rows, d_cols.shape[1]), dtype=np.float32)
    for r, s, e in zip(uniq, seg[:-1], seg[1:]):
        dw[r] = part(int(s), int(e), top)
    return dw
