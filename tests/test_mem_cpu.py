"""Host logic of the activation arena (_mem.py) and the graph cache bookkeeping (fused.GraphCache) on CPU tensors:
bump allocation, alignment, view reuse, growth between steps, lease ownership."""
import importlib

import torch

_mem = importlib.import_module("aread-multi-domain-recommendation_b200._mem")
fused = importlib.import_module("aread-multi-domain-recommendation_b200.fused")
CPU = torch.device("cpu")


def test_first_step_falls_back_then_arena_serves_everything():
    arena = _mem.Arena(CPU)
    lease = arena.acquire()
    assert lease is not None and arena.busy and arena.cap == 0
    with _mem.use(arena):
        a = _mem.empty((3, 5), torch.float32, CPU)          # no buffer yet: ordinary allocations, sizes recorded
        b = _mem.zeros((7,), torch.int32, CPU)
    assert arena.need == 256 + 256 and not b.any()
    lease.release()
    assert not arena.busy
    lease = arena.acquire()                                   # grows to 1.5 x the high-water mark
    assert arena.cap >= arena.need and arena.cap % 256 == 0
    with _mem.use(arena):
        a = _mem.empty((3, 5), torch.float32, CPU)
        b = _mem.zeros((7,), torch.int32, CPU)
    base = arena.buf.data_ptr()
    assert a.data_ptr() == base and b.data_ptr() == base + 256 and not b.any()
    assert a.shape == (3, 5) and a.dtype == torch.float32 and b.dtype == torch.int32
    lease.release()
    lease = arena.acquire()                                   # same request sequence: the very same view objects
    with _mem.use(arena):
        assert _mem.empty((3, 5), torch.float32, CPU) is a
        assert _mem.empty((7,), torch.int32, CPU) is b        # (zeros() would have cleared it)
        c = _mem.empty((1000, 1000), torch.float32, CPU)      # larger than ever seen: falls back, raises the mark
    assert c.data_ptr() < base or c.data_ptr() >= base + arena.cap
    assert arena.need > arena.cap
    cap = arena.cap
    lease.release()
    arena.acquire().release()
    assert arena.cap > cap                                    # grown between steps, never under a live lease


def test_lease_rules_and_routing():
    arena = _mem.Arena(CPU)
    first = arena.acquire()
    assert arena.acquire() is None                            # a forward that still awaits its backward owns it
    first.release()
    first.release()                                           # idempotent
    second = arena.acquire()
    first.release()                                           # a stale lease cannot free the current one
    assert arena.busy
    del second                                                # dropping the owner (autograd context) frees it
    assert not arena.busy
    t = _mem.empty((4,), torch.float32, CPU)                  # outside use(): plain torch
    assert t.shape == (4,)
    with _mem.use(None):
        assert _mem.zeros((2, 2), torch.float32, CPU).sum() == 0
    other = _mem.Arena(torch.device("meta"))
    with _mem.use(other):                                     # an arena of another device is ignored
        assert _mem.empty((4,), torch.float32, CPU).device.type == "cpu"


def test_disabled_arena(monkeypatch):
    monkeypatch.setattr(_mem, "ENABLED", False)
    assert _mem.Arena(CPU).acquire() is None


def test_graph_cache_bookkeeping():
    cache = fused.GraphCache()
    e = cache.get(("mask", 1))
    assert cache.get(("mask", 1)) is e and e.calls == 0 and e.fwd is None
    for i in range(fused.MAX_GRAPHS + 5):
        cache.get(("m", i))
    assert len(cache.entries) == fused.MAX_GRAPHS and ("mask", 1) not in cache.entries     # oldest evicted first
    like = torch.zeros(10, 4)
    g = cache.table_grad(like)
    assert cache.table_grad(like) is g and g.shape == like.shape
    assert cache.table_grad(torch.zeros(12, 4)) is not g      # a resized table gets a new buffer, the old one goes
    assert len(cache.tgrad) == 1
    cache.clear()
    assert not cache.entries and not cache.tgrad
