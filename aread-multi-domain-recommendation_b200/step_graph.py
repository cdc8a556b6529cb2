"""The whole train step of run.py:663-682 -- forward('domain_mask_bagging') + bagging BCE + L2 value + zero_grad +
backward + Adam -- as ONE CUDA graph per (domain mask, batch shape).

At the reference's own batch sizes (main.py:22 bs = 1024) a step is ~150 kernels of a few microseconds each: the host
cannot issue them as fast as the GPU retires them, whatever the kernels do.  Recording the whole step removes the
host from the loop: per step the trainer pays two small host->device copies (ids, labels), one for the dropout seed,
and one graph launch.

    step = GraphedTrainStep(model, optimizer)          # optimizer: optim.FusedAdam with the L2 term folded in
    loss = step(x, y, d)                               # device scalar (a static buffer: read it before the next call)

What makes it possible: the activation arena (every intermediate has a repeating address), the dropout seed and the
Adam step counters living in device memory, gradients written into the capture's private pool.  The first call for a
(mask, shape) runs eagerly (it sizes the arena and loads every kernel), the second one records.  Results are
bit-identical to the eager step (tests/test_step_graph_gpu.py).  Multi-GPU steps keep the per-mask forward / backward
sequences of fused.py (their collectives stay outside the recorded part).
"""
import torch

from . import _lib
from . import fused
from . import layer
from .optim import FusedAdam


class _Entry:
    __slots__ = ("graph", "x", "y", "loss", "seed", "plans", "calls", "n_launch")

    def __init__(self):
        self.graph = self.x = self.y = self.loss = self.seed = None
        self.plans, self.calls, self.n_launch = [], 0, 0


class GraphedTrainStep:
    def __init__(self, model, optimizer):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs optim.FusedAdam (its launch is capture-safe)")
        if not getattr(model, "_reg_folded", False):
            raise RuntimeError("fold the L2 gradient into the optimizer first: model.fold_regularization_into(optimizer)")
        self.model, self.opt = model, optimizer
        self.entries = {}
        self.pool = None

    def _eager(self, x, y, d):
        model = self.model
        preds = model(x, mode="domain_mask_bagging", domain_i=d)
        loss = model.bagging_loss(preds, y) + model.get_regularization_loss(device=x.device)
        model.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.detach().reshape(())

    def __call__(self, x, y, d):
        model = self.model
        info = model.mask_info(model.domain_mask[d])
        key = (info.serial, tuple(x.shape), tuple(y.shape), x.dtype, y.dtype, model.training)
        e = self.entries.get(key)
        if e is None:
            e = self.entries[key] = _Entry()
        e.calls += 1
        arena = model.arena(x.device)
        if e.graph is None and (e.calls < 2 or arena.busy or arena.need > arena.cap or not fused.USE_GRAPHS):
            return self._eager(x, y, d)
        if e.graph is None:
            self._record(e, x, y, d)
        e.x.copy_(x, non_blocking=True)
        e.y.copy_(y, non_blocking=True)
        if model.training and model.dropout_p > 0:
            host = torch.empty(1, dtype=torch.int64, pin_memory=True)
            host[0] = int(torch.randint(0, 2 ** 62, (1,)).item())
            e.seed.copy_(host, non_blocking=True)
        e.graph.replay()
        _lib.load().aread_launch_count_add(e.n_launch)          # the library's launches inside the recorded step
        self.opt.replayed(e.plans)
        model.embedding.plan(x.device).post_lookup(layer.BOUNDS_MODE)
        return e.loss

    def _record(self, e, x, y, d):
        model, opt = self.model, self.opt
        dev = x.device
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        e.x, e.y = x.clone(), y.clone()
        e.seed = torch.zeros(1, dtype=torch.int64, device=dev)
        opt.prepare_capture(dev)                                # host -> device copies happen before the capture starts
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = len(opt.captured_plans)
        launches0 = _lib.launch_count()
        use_graphs, bounds, seed_ptr = fused.USE_GRAPHS, layer.BOUNDS_MODE, fused.STEP_SEED_PTR
        fused.USE_GRAPHS, layer.BOUNDS_MODE, fused.STEP_SEED_PTR = False, "off", e.seed.data_ptr()
        rng = torch.get_rng_state()         # the recording draws a (discarded) seed: replays must see the caller's stream
        try:
            model.zero_grad()
            with fused._recording(graph, self.pool, dev):
                preds = model(e.x, mode="domain_mask_bagging", domain_i=d)
                loss = model.bagging_loss(preds, e.y) + model.get_regularization_loss(device=dev)
                loss.backward()
                opt.step()
                e.loss = loss.detach().reshape(())
        finally:
            fused.USE_GRAPHS, layer.BOUNDS_MODE, fused.STEP_SEED_PTR = use_graphs, bounds, seed_ptr
            torch.set_rng_state(rng)
        e.n_launch = _lib.launch_count() - launches0
        _lib.load().aread_launch_count_add(-e.n_launch & 0xFFFFFFFFFFFFFFFF)   # recorded, not run: replays count them
        e.plans = opt.captured_plans[n0:]
        del opt.captured_plans[n0:]
        e.graph = graph
        model.zero_grad()               # the recorded gradients live in the capture's pool; nothing outside needs them
