"""Host logic of the multi-GPU path (sharding.py) on CPU: the row -> (owner, local row) map, the
owner-major gradient layout, and the flat-bucket gradient all-reduce under a world_size-2 gloo group."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sharding = importlib.import_module("aread-multi-domain-recommendation_b200.sharding")


@pytest.mark.parametrize("n_rows,world", [(10, 2), (11, 2), (1, 2), (37, 4), (64, 8), (5, 8)])
def test_split_merge_round_trip(n_rows, world):
    full = torch.arange(n_rows * 3, dtype=torch.float32).reshape(n_rows, 3) + 1
    parts = [sharding.split_table(full, world, r) for r in range(world)]
    rows = sharding.shard_rows(n_rows, world)
    assert all(p.shape == (rows, 3) for p in parts)
    for r, p in enumerate(parts):
        for local in range(rows):
            g = local * world + r                        # the kernel's map: owner = row % world, local = row // world
            if g < n_rows:
                assert torch.equal(p[local], full[g])
            else:
                assert not p[local].any()                # padding rows are zero
    assert torch.equal(sharding.merge_shards(parts, n_rows), full)


@pytest.mark.parametrize("n_rows,world", [(11, 2), (37, 4)])
def test_owner_major_layout(n_rows, world):
    idx = sharding.owner_major_index(n_rows, world)
    rows = sharding.shard_rows(n_rows, world)
    assert idx.unique().numel() == n_rows and int(idx.max()) < world * rows
    g_full = torch.randn(n_rows, 4)
    buf = torch.zeros(world * rows, 4)
    buf[idx] = g_full
    for r in range(world):                               # rank r's slice of the buffer == its split of the gradient
        assert torch.equal(buf[r * rows:(r + 1) * rows], sharding.split_table(g_full, world, r))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) flat-bucket all-reduce == per-tensor average; a gradient that is None on one rank counts as zero
        torch.manual_seed(7)
        shapes = [(3, 5), (7,), (1, 1), (2, 3, 4)]
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(4))]
        all_grads = [[torch.randn(s) for s in shapes] for _ in range(world)]
        for i, (p, g) in enumerate(zip(params, all_grads[rank])):
            if not (rank == 1 and i == 2):               # rank 1 did not touch parameter 2 (a masked-out tower)
                p.grad = g.clone()
        sharding.allreduce_dense_grads(params)
        for i, p in enumerate(params[:-1]):
            want = sum(all_grads[r][i] for r in range(world) if not (r == 1 and i == 2)) / world
            torch.testing.assert_close(p.grad, want, rtol=1e-6, atol=1e-7)
        assert not params[-1].grad.any()                 # untouched everywhere: zero, present on every rank

        # (2) sharded table gradient: every rank scatters ITS batch into the owner-major layout, the
        # buffers are summed and averaged (the GPU path's reduce-scatter), and rank r's slice must be the
        # split of the batch-averaged dense gradient of the whole table
        dims, seq = np.array([13, 5, 9]), 3
        offsets = np.concatenate(([0], np.cumsum(dims)[:-1]))      # two one-hot columns + two history fields
        n_rows, D = int(dims.sum()), 4
        rng = np.random.default_rng(100 + rank)
        B = 6
        n_cols = 2 + 2 * seq
        x = np.zeros((B, n_cols), dtype=np.int64)
        x[:, 0] = rng.integers(0, 13, B)
        x[:, 1] = rng.integers(0, 5, B)
        x[:, 2:] = rng.integers(0, 9, (B, 2 * seq))
        d_out = rng.standard_normal((B, 4, D)).astype(np.float32)
        # plain restatement (kept independent of the oracle's argument conventions)
        g = np.zeros((n_rows, D), dtype=np.float64)
        for b in range(B):
            g[x[b, 0] + offsets[0]] += d_out[b, 0]
            g[x[b, 1] + offsets[1]] += d_out[b, 1]
            for f in range(2):
                for s_ in range(seq):
                    g[x[b, 2 + f * seq + s_] + offsets[2]] += d_out[b, 2 + f] / seq
        g = torch.from_numpy(g.astype(np.float32))
        rows = sharding.shard_rows(n_rows, world)
        buf = torch.zeros(world * rows, D)
        buf[sharding.owner_major_index(n_rows, world)] = g
        dist.all_reduce(buf)
        mine = buf[rank * rows:(rank + 1) * rows] / world
        g_avg = g.clone()
        dist.all_reduce(g_avg)
        g_avg /= world
        torch.testing.assert_close(mine, sharding.split_table(g_avg, world, rank), rtol=1e-6, atol=1e-7)

        # (3) checkpoint round trip: gather every rank's shard, merge, compare with the original table
        torch.manual_seed(3)
        full = torch.randn(n_rows, D)
        shard = sharding.split_table(full, world, rank)
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard)
        assert torch.equal(sharding.merge_shards(parts, n_rows), full)
        open(os.path.join(result_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
