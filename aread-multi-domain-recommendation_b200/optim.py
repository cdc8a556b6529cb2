"""FusedAdam: torch.optim.Adam's arithmetic (coupled weight decay, no amsgrad) as ONE kernel launch over
all parameter tensors (csrc/adam.cu).  Same constructor signature and state_dict layout as
torch.optim.Adam, so it can replace `torch.optim.Adam(model.parameters(), lr=..., betas=..., eps=...,
weight_decay=...)` at run.py:830 / 632 without other changes.  Parameters whose `.grad` is None are
skipped entirely (moments, step count and weights untouched), exactly like torch."""
import ctypes

import numpy as np
import torch

from . import _lib


class _Plan:
    """Launch table of one (param group, set of parameters that have a gradient): everything that does not
    change from step to step, laid out as the rows of the pinned block the kernel reads."""

    def __init__(self, params, idx, states, chunk, l2):
        self.params = params
        self.idx = np.asarray(idx, dtype=np.int64)
        self.ptrs = [p.data_ptr() for p in params]
        n = len(params)
        self.static = np.zeros((9, n), dtype=np.int64)
        h = self.static
        f = h.view(np.float32).reshape(9, -1)
        self.has_l2 = False
        start = 0
        for i, (p, st) in enumerate(zip(params, states)):
            h[0, i] = p.data_ptr()
            h[2, i], h[3, i] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            h[4, i], h[5, i] = p.numel(), start
            start += (p.numel() + chunk - 1) // chunk
            c = l2.get(p, 0.0)
            if c > 0:
                f[8, i] = np.float32(2.0) * np.float32(c)
                self.has_l2 = True
        self.n_chunks = start
        self.moments = [(st["exp_avg"], st["exp_avg_sq"]) for st in states]     # keeps the addresses valid
        # CUDA-graph capture (step_graph.py): a pinned copy of the launch table whose content never changes after the
        # capture (the recorded upload reads it at every replay) and the positions of the step counters
        self.graph_host = self.graph_dev = self.slot_dev = None


class MultiCopy:
    """dst[i] <- src[i] (or zeros when `srcs` is None) for a fixed list of same-device CUDA tensors in ONE launch
    (csrc/adam.cu multi_copy_kernel).  The launch table is built once and kept on the device."""

    def __init__(self, dsts, srcs=None):
        lib = _lib.load()
        chunk = int(lib.aread_multi_copy_chunk())
        self.dsts, self.srcs = list(dsts), None if srcs is None else list(srcs)
        n = len(self.dsts)
        self.n = n
        if n == 0:
            return
        dev = self.dsts[0].device
        table = np.zeros((4, n), dtype=np.int64)
        start = 0
        for i, d in enumerate(self.dsts):
            if not (d.is_cuda and d.is_contiguous() and d.device == dev):
                raise RuntimeError("MultiCopy handles contiguous CUDA tensors of one device")
            nbytes = d.numel() * d.element_size()
            if nbytes % 4 or d.data_ptr() % 4:
                raise RuntimeError("MultiCopy needs 4-byte multiples")
            table[0, i] = d.data_ptr()
            if self.srcs is not None:
                src = self.srcs[i]
                if src.dtype != d.dtype or src.shape != d.shape or not src.is_contiguous() or src.device != dev:
                    raise RuntimeError("MultiCopy: source / destination mismatch")
                table[1, i] = src.data_ptr()
            table[2, i], table[3, i] = nbytes, start
            start += (nbytes + chunk - 1) // chunk
        self.n_chunks = start
        self.ptrs = table[:2].copy()
        self.table = torch.from_numpy(table).to(dev)
        self.device = dev

    def current(self):
        """False once one of the tensors moved (the table holds raw addresses)."""
        return self.n == 0 or (all(int(self.ptrs[0, i]) == d.data_ptr() for i, d in enumerate(self.dsts)) and
                               (self.srcs is None or
                                all(int(self.ptrs[1, i]) == t.data_ptr() for i, t in enumerate(self.srcs))))

    def run(self):
        if self.n == 0:
            return
        base, row = self.table.data_ptr(), self.table.stride(0) * 8
        args = _lib.MultiCopyArgs(self.n, self.n_chunks, base, base + row if self.srcs is not None else None,
                                  base + 2 * row, base + 3 * row)
        _lib.check(_lib.load().aread_multi_copy(ctypes.byref(args),
                                                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._chunk = None
        self._l2 = {}
        self._bind_steps()

    def fold_l2(self, terms):
        """terms: {parameter: l2}.  The step then uses grad + 2 * l2 * p for these parameters (a missing
        gradient counts as zero: a regularised parameter is updated every step, exactly as when the
        regulariser is part of the loss)."""
        known = {p for group in self.param_groups for p in group["params"]}
        missing = [p for p in terms if p not in known]
        if missing:
            raise ValueError(f"{len(missing)} regularised parameter(s) are not managed by this optimizer")
        self._l2 = dict(terms)
        self._plans = {}

    def reset(self):
        """Back to the state of a freshly constructed optimizer -- step counts 0, both moments 0 -- WITHOUT
        reallocating the moments: what the trainer's per-candidate `optimizer_fast = Adam(model.parameters(), ...)`
        (run.py:632-633) amounts to, as one zero-fill launch instead of a table-sized allocation per candidate.
        Hyper-parameters are kept; change them through `param_groups` as usual."""
        moments = []
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    moments += [st["exp_avg"], st["exp_avg_sq"]]
        for arr, _ in self._steps:
            arr[:] = 0
        for ctr in self._counters.values():
            ctr.zero_()
        if moments:
            z = getattr(self, "_zero", None)
            if z is None or len(z.dsts) != len(moments) or any(a is not b for a, b in zip(z.dsts, moments)) \
                    or not z.current():
                z = self._zero = MultiCopy(moments)
            z.run()

    def _device_counters(self, gi, device):
        """fp32 step counters of group gi on the device (graph replay derives the bias corrections from them); kept
        equal to the host counters: uploaded here, then advanced on both sides by every captured / replayed step."""
        key = (gi, device)
        ctr = self._counters.get(key)
        if ctr is None:
            ctr = self._counters[key] = torch.from_numpy(self._steps[gi][0].copy()).to(device)
        return ctr

    def sync_counters(self):
        """Host step counts -> device counters (after eager steps, load_state_dict or reset)."""
        for (gi, device), ctr in self._counters.items():
            ctr.copy_(torch.from_numpy(self._steps[gi][0].copy()), non_blocking=False)

    def prepare_capture(self, device):
        """Everything a recorded step needs that is a host->device copy: the device step counters (equal to the host
        counts) and, per known launch plan, the positions of its tensors' counters."""
        self.sync_counters()
        for gi in range(len(self.param_groups)):
            self._device_counters(gi, device)
        for plan in self._plans.values():
            if plan.slot_dev is None:
                plan.slot_dev = torch.from_numpy(plan.idx.copy()).to(device)

    def replayed(self, records):
        """Book-keeping of a replayed whole-step graph: the host mirror of the step counters."""
        for rec in records:
            self._steps[rec[0]][0][rec[1].idx] += 1

    def _bind_steps(self):
        """state[p]["step"] (a host scalar tensor, as in torch.optim.Adam) becomes a view into one array per
        group, so that a step bumps all counters with one vector add."""
        self._steps, self._plans = [], {}
        self._counters = getattr(self, "_counters", {})
        self._counters.clear()
        self.captured_plans = []
        for group in self.param_groups:
            arr = np.zeros(len(group["params"]), dtype=np.float32)
            view = torch.from_numpy(arr)
            for i, p in enumerate(group["params"]):
                st = self.state.get(p)
                if st:
                    arr[i] = float(st["step"])
                    st["step"] = view[i]
            self._steps.append((arr, view))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._bind_steps()          # (also drops the device counters: whole-step graphs must be recorded again)

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_steps"):
            self._bind_steps()

    def _plan(self, gi, group, idx):
        params = [group["params"][i] for i in idx]
        arr, view = self._steps[gi]
        states = []
        for i, p in zip(idx, params):
            if (p.grad is not None and p.grad.is_sparse) or not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FusedAdam handles dense fp32 CUDA parameters only")
            st = self.state[p]
            if not st:
                st["step"] = view[i]                                            # host scalar, like torch's default
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            states.append(st)
        return _Plan(params, idx, states, self._chunk, self._l2)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if self._chunk is None:
            self._chunk = int(lib.aread_adam_chunk())
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group["betas"]
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            grads = [p.grad for p in group["params"]]
            if self._l2:    # a folded regulariser gives its parameters a gradient every step
                l2 = self._l2
                idx = tuple(i for i, (g, p) in enumerate(zip(grads, group["params"])) if g is not None or p in l2)
            else:
                idx = tuple(i for i, g in enumerate(grads) if g is not None)    # parameters without gradient: skipped
            if not idx:
                continue
            plan = self._plans.get((gi, idx))
            if plan is None or plan.ptrs != [p.data_ptr() for p in plan.params]:
                if len(self._plans) > 256:
                    self._plans.clear()
                plan = self._plans[(gi, idx)] = self._plan(gi, group, idx)
            grads = [g if (g is None or g.is_contiguous()) else g.contiguous() for g in (grads[i] for i in idx)]
            n = len(idx)
            device = plan.params[0].device
            steps = self._steps[gi][0]
            if torch.cuda.is_current_stream_capturing():
                # whole-step CUDA graph: nothing of this launch may change between replays.  The table (with the
                # gradient addresses of THIS capture) is uploaded from a pinned block that is never written again, the
                # bias corrections come from device-side step counters
                ctr = self._device_counters(gi, device)
                host = torch.empty((9, n), dtype=torch.int64, pin_memory=True)
                h = host.numpy()
                np.copyto(h, plan.static)
                h[1, :] = [0 if g is None else g.data_ptr() for g in grads]
                if plan.slot_dev is None:
                    raise RuntimeError("FusedAdam: call prepare_capture() before recording a step")
                dev_table = torch.empty((9, n), dtype=torch.int64, device=device)
                dev_table.copy_(host, non_blocking=True)
                base, row = dev_table.data_ptr(), dev_table.stride(0) * 8
                args = _lib.AdamArgs(n, plan.n_chunks, base, base + row, base + 2 * row, base + 3 * row, base + 4 * row,
                                     base + 5 * row, None, None, beta1, beta2, eps, wd,
                                     base + 8 * row if plan.has_l2 else None, ctr.data_ptr(), plan.slot_dev.data_ptr(), lr,
                                     beta1, beta2)
                _lib.check(lib.aread_adam_step(ctypes.byref(args),
                                               ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
                # the recorded upload re-reads `host` at every replay: it lives as long as the graph that owns this record
                self.captured_plans.append((gi, plan, host, dev_table, grads))
                continue
            steps[plan.idx] += 1
            t = steps[plan.idx].astype(np.float64)
            # rows: params, grads, exp_avg, exp_avg_sq, sizes, chunk_start (int64); step_size, bc2_sqrt, 2 * l2 (fp32).
            # A fresh pinned block per step: torch's host allocator will not recycle it before the copy ran.
            host = torch.empty((9, n), dtype=torch.int64, pin_memory=True)
            dev = torch.empty((9, n), dtype=torch.int64, device=device)
            h = host.numpy()
            np.copyto(h, plan.static)
            h[1, :] = [0 if g is None else g.data_ptr() for g in grads]
            f = h.view(np.float32).reshape(9, -1)                               # fp32 view of the same rows
            f[6, :n] = lr / (1.0 - beta1 ** t)
            f[7, :n] = np.sqrt(1.0 - beta2 ** t)
            dev.copy_(host, non_blocking=True)
            base, row = dev.data_ptr(), dev.stride(0) * 8
            args = _lib.AdamArgs(n, plan.n_chunks, base, base + row, base + 2 * row, base + 3 * row, base + 4 * row,
                                 base + 5 * row, base + 6 * row, base + 7 * row, beta1, beta2, eps, wd,
                                 base + 8 * row if plan.has_l2 else None, None, None, lr, beta1, beta2)
            _lib.check(lib.aread_adam_step(ctypes.byref(args),
                                           ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
            ctr = self._counters.get((gi, device))
            if ctr is not None:         # an eager step between graph replays: keep the device counters in step
                ctr.index_add_(0, torch.from_numpy(plan.idx).to(device), torch.ones(n, device=device))
        return loss
