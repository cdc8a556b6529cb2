"""Row-sharded embedding table over the GPUs of one NVSwitch box (one process per GPU).

Global row r lives on rank r mod N at local row r // N (N a power of two: balances the skewed big
fields).  Every rank maps all peers' shards into its own address space (CUDA IPC through torch's
tensor-sharing plumbing) and the gather kernel reads remote rows straight over NVLink -- the lookup
is the all-to-all of rows, no index exchange, no staging.  The gradient goes the other way the same way: every
rank keeps one receive buffer per sender ([N, rows_per_shard, D], zero between steps), maps its peers' buffers, and the
scatter kernel stores the reduced gradient row of table row r straight into slot `sender` of the rank that owns r
(P2P stores over NVLink).  Only the rows a batch touched travel -- O(unique rows x D) instead of the O(R x D) dense
reduce-scatter of a full-size buffer -- and the owner adds its N slots in rank order (deterministic), clearing what it
read.  Dense parameters are all-reduced in one flat bucket.  AREAD_DENSE_GRAD_EXCHANGE=1 keeps the dense NCCL
reduce-scatter (the baseline the sparse exchange is measured against).
"""
import ctypes
import math
import os

import torch
import torch.distributed as dist

from . import _lib


def shard_rows(n_rows, world):
    return (n_rows + world - 1) // world


def split_table(full, world, rank):
    """Rows rank, rank + world, ... of `full` ([R, D]) padded with zero rows to shard_rows(R, world)."""
    n = shard_rows(full.shape[0], world)
    out = torch.zeros((n, full.shape[1]), dtype=full.dtype, device=full.device)
    mine = full[rank::world]
    out[:mine.shape[0]] = mine
    return out


def owner_major_index(n_rows, world):
    """Position of every global row in the owner-major [world * shard_rows, D] gradient buffer the
    scatter kernel writes (embedding.cu dest_row): owner * shard_rows + local row."""
    r = torch.arange(n_rows, dtype=torch.int64)
    return (r % world) * shard_rows(n_rows, world) + r // world


def merge_shards(shards, n_rows):
    """Inverse of split_table for a list of all ranks' shards (checkpointing / tests)."""
    world = len(shards)
    full = torch.empty((n_rows, shards[0].shape[1]), dtype=shards[0].dtype, device=shards[0].device)
    for r, s in enumerate(shards):
        full[r::world] = s[:full[r::world].shape[0]]
    return full


class TableShards:
    """Peer-mapped views of every rank's shard + the collectives of the sharded lookup."""

    def __init__(self, shard_param, n_rows, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world & (self.world - 1):
            raise ValueError("row sharding needs a power-of-two number of GPUs")
        self.shift = int(math.log2(self.world))
        self.n_rows = n_rows
        self.rows = shard_param.shape[0]
        self.dim = shard_param.shape[1]
        dev = shard_param.device
        lib = _lib.load()
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_int64(0)
        _lib.check(lib.aread_ipc_export(ctypes.c_void_p(shard_param.data_ptr()), handle, ctypes.byref(offset)))
        metas = [None] * self.world
        dist.all_gather_object(metas, (dev.index, handle.raw, int(offset.value)), group=group)
        self._shard = shard_param                          # the exported allocation must stay alive
        self._opened = []
        ptrs = []
        for r, (peer_index, raw, off) in enumerate(metas):
            if r == self.rank:
                ptrs.append(shard_param.data_ptr())
                continue
            if not torch.cuda.can_device_access_peer(dev.index, peer_index):
                raise RuntimeError(f"GPU {dev.index} cannot access GPU {peer_index} (no NVLink/P2P path)")
            out = ctypes.c_void_p(0)
            _lib.check(lib.aread_ipc_open(raw, off, dev.index, ctypes.byref(out)))
            self._opened.append((out.value, off))
            ptrs.append(out.value)
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self._token = torch.zeros(1, dtype=torch.float32, device=dev)
        self._grad = None
        # sparse gradient exchange: this rank's receive slots, mapped by every peer
        self.sparse = os.environ.get("AREAD_DENSE_GRAD_EXCHANGE", "0") != "1"
        self.recv = self.push_ptrs = None
        if self.sparse:
            self.recv = torch.zeros((self.world, self.rows, self.dim), dtype=torch.float32, device=dev)
            _lib.check(lib.aread_ipc_export(ctypes.c_void_p(self.recv.data_ptr()), handle, ctypes.byref(offset)))
            metas = [None] * self.world
            dist.all_gather_object(metas, (dev.index, handle.raw, int(offset.value)), group=group)
            slot = self.rank * self.rows * self.dim * 4              # this sender's slot inside every owner's buffer
            push = []
            for r, (peer_index, raw, off) in enumerate(metas):
                if r == self.rank:
                    push.append(self.recv.data_ptr() + slot)
                    continue
                out = ctypes.c_void_p(0)
                _lib.check(lib.aread_ipc_open(raw, off, dev.index, ctypes.byref(out)))
                self._opened.append((out.value, off))
                push.append(out.value + slot)
            self.push_ptrs = torch.tensor(push, dtype=torch.int64, device=dev)
        dist.barrier(group=group)

    def close(self):
        lib = _lib.load()
        for ptr, off in self._opened:
            lib.aread_ipc_close(ctypes.c_void_p(ptr), off)
        self._opened = []

    def fence(self):
        """Stream-ordered barrier: peers may read this rank's shard only after its optimizer step, and
        it may read theirs only after theirs."""
        dist.all_reduce(self._token, group=self.group)

    def grad_buffer(self, device):
        if self._grad is None:
            self._grad = torch.empty((self.world * self.rows, self.dim), dtype=torch.float32, device=device)
        return self._grad

    def reduce_grad(self, d_full):
        """-> batch-averaged gradient of this rank's shard.  Dense exchange: `d_full` is the owner-major
        [world * rows, D] local gradient, reduce-scattered by NCCL.  Sparse exchange: the scatter kernels have already
        stored every sender's reduced rows into this rank's receive slots; once all ranks have reached this point
        (fence) the slots are added in rank order and cleared."""
        out = torch.empty((self.rows, self.dim), dtype=torch.float32, device=self.recv.device if self.sparse else d_full.device)
        if not self.sparse:
            dist.reduce_scatter_tensor(out, d_full, op=dist.ReduceOp.AVG, group=self.group)
            return out
        self.fence()
        stream = ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
        _lib.check(_lib.load().aread_shard_grad_sum(ctypes.c_void_p(self.recv.data_ptr()), self.world,
                                                    self.rows * self.dim, 1.0 / self.world,
                                                    ctypes.c_void_p(out.data_ptr()), stream))
        return out


def allreduce_dense_grads(params, group=None):
    """One flat-bucket all-reduce (average) of the gradients of the replicated parameters.

    EVERY parameter takes part on every rank: ranks train different domains, so a tower that is masked
    out on one rank (gradient None) is live on another; a missing gradient counts as zero, like DDP's
    unused parameters.  Afterwards each p.grad is a view into the reduced bucket."""
    params = list(params)
    if not params:
        return
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=params[0].dtype, device=params[0].device)
    views = [c.view_as(p) for c, p in zip(flat.split(sizes), params)]
    have = [(v, p.grad) for v, p in zip(views, params) if p.grad is not None]
    if have:
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:                                  # gloo (the CPU tests of this host logic) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
    for p, v in zip(params, views):
        p.grad = v
