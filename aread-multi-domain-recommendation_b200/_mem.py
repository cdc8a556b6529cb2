"""Per-model activation arena of the fused train step.

Every intermediate of the fused node (fused.py) -- looked-up rows, expert / tower activations, the
backward's temporaries -- is carved out of ONE persistent device buffer with a bump pointer instead of
going through the caching allocator.  The allocator's behaviour depends on the HEMP mask (compact
activations change size from domain to domain), which on a multi-GPU box turned into cudaMalloc calls
in the middle of steps (each maps the new block into every peer: ~2 ms); the arena makes a step
allocation-free and its addresses reproducible, which is also what CUDA-graph replay needs.

Only intermediates live here.  What leaves the node -- probabilities, recorded gates, parameter
gradients -- is allocated normally, so nothing the caller holds is ever overwritten.

Ownership: a forward `acquire`s the arena and holds the lease until its backward has run (or its
autograd context is dropped).  While a lease is out, further forwards of the same model fall back to
normal allocation -- correct, just not allocation-free.  A second backward through the same context
(retain_graph=True) would read recycled buffers, so it is refused.
"""
import os

import torch

ENABLED = os.environ.get("AREAD_WORKSPACE", "1") != "0"
_ALIGN = 256
_GROW = 1.5
_current = None           # the arena the running forward / backward allocates from (None: torch.empty)


class Lease:
    __slots__ = ("arena", "gen", "active")

    def __init__(self, arena, gen):
        self.arena, self.gen, self.active = arena, gen, True

    def release(self):
        if self.active:
            self.active = False
            if self.arena.gen == self.gen:
                self.arena.busy = False

    def __del__(self):
        self.release()


class Arena:
    def __init__(self, device):
        self.device = device
        self.buf = None
        self.cap = 0
        self.off = 0
        self.need = 0          # high-water mark of any step so far (bytes)
        self.gen = 0
        self.buf_gen = 0       # bumped whenever `buf` is reallocated (recorded CUDA graphs key on it)
        self.busy = False
        self.views = {}

    def acquire(self):
        """Lease for one forward + backward, or None when another forward still owns the arena."""
        if self.busy or not ENABLED:
            return None
        if self.need > self.cap:                       # grow between steps, never under a live lease
            self.buf = None
            self.views.clear()
            self.cap = int(self.need * _GROW) // _ALIGN * _ALIGN
            self.buf = torch.empty(self.cap, dtype=torch.uint8, device=self.device)
            self.buf_gen += 1
        elif len(self.views) > 65536:
            self.views.clear()
        self.off = 0
        self.gen += 1
        self.busy = True
        return Lease(self, self.gen)

    def alloc(self, shape, dtype):
        n = dtype.itemsize
        for s in shape:
            n *= s
        start = self.off
        self.off = end = start + (n + _ALIGN - 1) // _ALIGN * _ALIGN
        if end > self.need:
            self.need = end
        if end > self.cap:                             # first steps / a larger mask than seen so far
            return torch.empty(shape, dtype=dtype, device=self.device)
        key = (start, shape, dtype)
        t = self.views.get(key)
        if t is None:
            t = self.buf[start:start + n].view(dtype).view(shape)
            self.views[key] = t
        return t


class use:
    """with use(arena_or_None): ...   -- routes empty()/zeros() below to the arena."""

    def __init__(self, arena):
        self.arena = arena

    def __enter__(self):
        global _current
        self.prev, _current = _current, self.arena

    def __exit__(self, *exc):
        global _current
        _current = self.prev


def empty(shape, dtype, device):
    a = _current
    if a is None or a.device != device:
        return torch.empty(shape, dtype=dtype, device=device)
    return a.alloc(tuple(shape), dtype)


def zeros(shape, dtype, device):
    a = _current
    if a is None or a.device != device:
        return torch.zeros(shape, dtype=dtype, device=device)
    return a.alloc(tuple(shape), dtype).zero_()


_WORKSPACES = {}


def workspace(name, device, need):
    """Scratch bytes for ONE kernel call (partial sums, sort buffers).  Inside the fused node they come from the
    activation arena like every other intermediate, so a recorded CUDA graph never points at a buffer that a later,
    larger call reallocates; outside it from a grow-only buffer per (name, device)."""
    need = max(int(need), 256)
    a = _current
    if a is not None and a.device == device:
        return a.alloc((need,), torch.uint8)
    key = (name, device)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < need:
        ws = _WORKSPACES[key] = torch.empty(need, dtype=torch.uint8, device=device)
    return ws
