"""Host-side mirror of the reference `model/aread.py`: the AREAD nn.Module.

The class keeps the reference's constructor / forward signature, attribute names, `state_dict`
layout and HEMP mask API so that `run.py` drives it unchanged.  The forward runs
on the C-ABI CUDA library through `dense_ops` / `embedding_ops`; the HEMP bookkeeping is host
logic (`hemp.py`).  Reference lines are cited per method.
"""
import copy
import os
import re

import numpy as np
import torch
from torch import nn

from . import _mem, dense_ops, fused, hemp, loss_ops
from .expert_ops import ExpertLayer
from .layer import BaseModel, CrossNetwork, MultiLayerPerceptron, _weights_without_bn
from .packing import PackSet


class AREAD(BaseModel):
    """Adaptive REcommendation for All Domains (reference: model/aread.py:15-680)."""

    def __init__(self, one_hot_feature_dims, embed_dim, multi_hot_dict, n_tower, n_domain, base_model,
                 expert_dims, tower_dims, domain_idx,
                 domain2group=None, n_cross_layers=3, dropout=0.2, device=None,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, config=None):
        super().__init__(one_hot_feature_dims, embed_dim, multi_hot_dict,
                         l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        if base_model != 'mmoe':
            # the 'ple' bottom cannot be built from the shipped config
            raise NotImplementedError("aread_b200 implements the 'mmoe' bottom of AREAD only")
        self.model_name = 'aread'
        self.base_model = base_model
        self.domain_idx = domain_idx
        self.n_tower = n_tower
        self.n_level = len(n_tower)
        self.edge_num = n_tower[0] + np.sum([n_tower[l - 1] * n_tower[l] for l in range(1, self.n_level)]) \
            + n_tower[-1]
        self.n_domain = n_domain
        self.tower_dims = tower_dims
        self.bottom_level = len(expert_dims)
        self.device = device
        self.dropout_p = dropout
        # operand precision of the expert GEMMs: 'bf16' (one tensor-core pass, BASELINE "bf16 experts") or
        # 'bf16x3' (split hi/lo operands, three passes, fp32-grade results)
        self.expert_precision = os.environ.get("AREAD_EXPERT_PRECISION", "bf16")
        if self.expert_precision not in ("bf16", "bf16x3"):
            raise ValueError(f"AREAD_EXPERT_PRECISION must be 'bf16' or 'bf16x3', got {self.expert_precision!r}")
        self.domain2group = None if domain2group is None else np.array([domain2group[d] for d in range(n_domain)])
        self.domain_mask = [None] * n_domain
        self.candidate_domain_mask = None
        self.tower2cluster = [[None] * n_tower[l] for l in range(self.n_level)]
        self.model_state = None
        self.domain_tower_gate_values = None
        self.tmp_tower_gate_values = [[None] * n_tower[l] for l in range(self.n_level)]
        self.gate_value_threshold = None
        self.eval_loss = None
        self.group_embedding = nn.Embedding(n_tower[0], embed_dim)
        self.final_gate = nn.Sequential(nn.Linear(2 * embed_dim, n_tower[-1], bias=False), nn.Softmax(dim=1))
        self.domain_size = np.array(config.domain_size[config.dataset_name])
        self.use_dcn = getattr(config, 'use_dcn', False)
        self.use_atten = getattr(config, 'use_atten', False)
        if not self.use_dcn:
            # the reference forward reads cn_out unconditionally in every masked mode (aread.py:231, 242)
            raise ValueError("AREAD needs config.use_dcn = True (the reference fails without it)")
        self.cn = CrossNetwork(self.embed_output_dim, config.n_cross_layers)
        if self.use_atten:
            # parameters only: the reference computes this branch and never reads the result
            # (aread.py:139-140), so it is kept for state_dict compatibility and never executed
            self.build_atten(config, dropout)

        n_expert = config.mmoe_n_expert
        self.mmoe_experts = nn.ModuleList(
            MultiLayerPerceptron(self.embed_output_dim, expert_dims, dropout, output_layer=False)
            for _ in range(n_expert))
        self.mmoe_gates = nn.ModuleList(
            nn.Sequential(nn.Linear(self.embed_output_dim, n_expert), nn.Softmax(dim=1)) for _ in range(n_tower[0]))
        self.add_regularization_weight(_weights_without_bn(self.mmoe_experts), l2=l2_reg_dnn)

        towers, gates = [], []
        width = expert_dims[-1]
        for l in range(self.n_level):
            towers.append(nn.ModuleList(MultiLayerPerceptron(width, tower_dims[l], dropout, output_layer=False)
                                        for _ in range(n_tower[l])))
            if l > 0:
                gates.append(nn.ModuleList(nn.Sequential(nn.Linear(2 * embed_dim, n_tower[l - 1]))
                                           for _ in range(n_tower[l])))
            width = tower_dims[l][-1]
        self.towers = nn.ModuleList(towers)
        self.tower_gates = nn.ModuleList(gates)
        self.towers_linear = nn.ModuleList(nn.Linear(self.embed_output_dim + width, 1, bias=False)
                                           for _ in range(n_tower[-1]))
        self.output_layers = nn.ModuleList(nn.Sigmoid() for _ in range(n_tower[-1]))
        self.add_regularization_weight(_weights_without_bn(self.towers), l2=l2_reg_dnn)
        self.add_regularization_weight(_weights_without_bn(self.cn), l2=l2_reg_cross)
        self._mask_cache = {}
        self._build_packs()

    # ----------------------------------------------------------------------------- packed storage
    def _build_packs(self):
        """Re-point the per-expert parameters at packed storage (packing.py) so that the grouped
        kernels see all experts of a layer side by side."""
        packs = PackSet()
        layers = []
        for i in range(self.bottom_level):
            layers.append(ExpertLayer(packs, [e.layers[4 * i] for e in self.mmoe_experts],
                                      [e.layers[4 * i + 1] for e in self.mmoe_experts], salt=0x1000 + i))
        tower_layers = []
        for l in range(self.n_level):
            per_level = []
            for j in range(len(self.tower_dims[l])):
                per_level.append(ExpertLayer(packs, [t.layers[4 * j] for t in self.towers[l]],
                                             [t.layers[4 * j + 1] for t in self.towers[l]],
                                             salt=0x2000 + 16 * l + j))
            tower_layers.append(per_level)
        object.__setattr__(self, "_packs", packs)
        object.__setattr__(self, "_expert_layers", layers)
        object.__setattr__(self, "_tower_layers", tower_layers)
        object.__setattr__(self, "_fused", fused.ModelPacks(self, packs, layers, tower_layers))
        object.__setattr__(self, "_fused_params", fused.param_list(self))
        object.__setattr__(self, "_slot_cache", {})
        object.__setattr__(self, "_arenas", {})
        object.__setattr__(self, "_graphs", fused.GraphCache())
        object.__setattr__(self, "_rollback", None)
        object.__setattr__(self, "_wcast", None)
        object.__setattr__(self, "_mixed_tables", {})

    @staticmethod
    def bagging_loss(y_stack, targets):
        """sum_t BCELoss(y_stack[t], targets) / n_act -- the trainer's loss of the 'domain_mask_bagging' output
        (run.py:643-644, 672-677) as one kernel."""
        return loss_ops.bagging_bce(y_stack, targets)

    def record_graphs(self, x, domains=None, mode="domain_mask_bagging", backward=True):
        """Record the CUDA-graph launch sequences (fused.py) for batches shaped like `x` under the masks of
        `domains` ahead of time -- e.g. once after `update_all_mask`.  Without this call they are recorded lazily
        after a few eager steps per mask.  Parameters, buffers, gradients and the CPU random stream are left as
        they were."""
        domains = range(self.n_domain) if domains is None else domains
        buffers = {n: b.detach().clone() for n, b in self.named_buffers()}
        had_grad = any(p.grad is not None for p in self.parameters())
        rng = torch.get_rng_state()
        # one pass per DISTINCT mask, largest first: the activation arena then has its final size before most
        # sequences are recorded.  Per mask: one eager pass (every kernel it uses has then been loaded), one
        # recording pass; a pass that finds the arena too small runs eagerly and is repeated
        infos = {}
        for d in domains:
            info = self.mask_info(self.domain_mask[d])
            infos.setdefault(info.serial, (d, info))
        order = sorted(infos.values(), key=lambda di: -sum(len(a) for a in di[1].active_idx))

        def one_pass(d):
            if backward and not had_grad and self.training:
                self(x, mode=mode, domain_i=d).sum().backward()
                self.zero_grad()
            else:
                with torch.no_grad():
                    self(x, mode=mode, domain_i=d)

        self._graphs.force = True
        try:
            for d, info in order:
                for _ in range(4):
                    one_pass(d)
                    if self._graphs.recorded(info.serial, x.shape, self.training):
                        break
        finally:
            self._graphs.force = False
        with torch.no_grad():
            for n, b in self.named_buffers():
                b.copy_(buffers[n])
        torch.set_rng_state(rng)

    def expert_weights_bf16(self):
        """bf16 operand copies of the expert Linear weights ([groups * n, k] per layer), refreshed with one launch."""
        flats = [L.weight.flat for L in self._expert_layers]
        wc = self._wcast
        if wc is None or wc.src_ptrs != [f.data_ptr() for f in flats]:
            from .dense_kernels import WeightCast
            wc = WeightCast(flats)
            object.__setattr__(self, "_wcast", wc)
        return [o.view(-1, o.shape[-1]) for o in wc.run()]

    def arena(self, device):
        """The activation arena of the fused step on `device` (_mem.py)."""
        a = self._arenas.get(device)
        if a is None:
            a = self._arenas[device] = _mem.Arena(device)
        return a

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_packs", None) is not None:
            for pack in self._packs.packs:          # .to() / .cuda() replaced every .data: pack again
                pack.repack()
            self._graphs.clear()                    # recorded launch sequences point at the old storage
        return out

    def __deepcopy__(self, memo):
        cls = self.__class__
        clone = cls.__new__(cls)
        memo[id(self)] = clone
        for k, v in self.__dict__.items():
            if k not in ("_packs", "_expert_layers", "_tower_layers", "_fused", "_fused_params", "_slot_cache", "_arenas", "_graphs",
                         "_rollback", "_wcast", "_mixed_tables"):
                setattr(clone, k, copy.deepcopy(v, memo))
        clone._build_packs()
        return clone

    # ------------------------------------------------------------------------------------ forward
    def forward(self, x, mode='wo_mask', targets=None, memory_gate_value=False,
                domain_i=None, current_mask=None, tmp_memory_gate_value=False):
        """Reference: aread.py:129-261.  Live modes: 'wo_mask', 'domain_with_mask',
        'domain_mask_bagging'; 'domain_mask_final' is kept callable (full masks only, as in the
        reference); 'with_mask' raises like the reference does (undefined name at aread.py:215)."""
        if mode == 'with_mask':
            raise NameError("name 'other_outs' is not defined")      # reference behaviour, aread.py:215
        if mode == 'wo_mask':
            out = dense_ops.aread_forward(self, x, None, want_gates=memory_gate_value)
            y = out.probs.mean(dim=0).unsqueeze(-1)                                       # [B, 1]
            if memory_gate_value:
                self._record_unmasked_gates(x, out.gates, domain_i)
            return y
        if mode not in ('domain_with_mask', 'domain_mask_bagging', 'domain_mask_final'):
            raise ValueError(f"unknown forward mode '{mode}'")
        mask = self.domain_mask[domain_i] if current_mask is None else current_mask
        info = self.mask_info(mask)
        if mode == 'domain_mask_final':
            with torch.no_grad():
                out = dense_ops.aread_forward(self, x, info, want_gate_means=memory_gate_value, want_gate_inputs=True)
            self._store_gate_means(out, info, domain_i, memory_gate_value, False)
            gate = self.final_gate(out.gate_inputs.detach()) * mask[-1].squeeze(1)
            gate = gate / (gate.sum(dim=1, keepdim=True) + 1e-8)
            return torch.sum(out.probs.transpose(0, 1) * gate, dim=1)
        # candidate masks of the HEMP search (current_mask=...) are scored a handful of times: not worth a CUDA graph
        out = dense_ops.aread_forward(self, x, info, want_gate_means=memory_gate_value or tmp_memory_gate_value,
                                      may_record=current_mask is None)
        self._store_gate_means(out, info, domain_i, memory_gate_value, tmp_memory_gate_value)
        return out.probs if mode == 'domain_mask_bagging' else out.probs.mean(dim=0)

    def forward_mixed(self, x, return_stack=False):
        """Eval-mode 'domain_with_mask' for a batch that MIXES domains: row b is evaluated under
        `domain_mask[x[b, domain_idx]]`, all rows in one launch sequence (mixed_ops.py, csrc/mixed.cu).  Equals calling
        `forward(x_d, mode='domain_with_mask', domain_i=d)` once per domain (run.py:719-727) row for row; rows sorted
        by domain let the kernel skip the (tile, tower) pairs their masks prune.  Returns y [B] (and, with
        return_stack, y_stack [n_tower[-1], B] with zeros for pruned heads)."""
        from . import mixed_ops
        return mixed_ops.forward_mixed_eval(self, x, want_stack=return_stack)

    def hier_tower_mask_forward(self, d, tower_inputs, gate_inputs, domain_cn_out, domain_linear_out,
                                single_domain_mask, memory_gate_value=False, tmp_memory_gate_value=False):
        """Reference: aread.py:263-322.  Kept for API completeness; `forward` does not go through it."""
        info = self.mask_info(single_domain_mask)
        out = dense_ops.hei_forward(self, tower_inputs, gate_inputs, domain_cn_out, domain_linear_out, info,
                                    want_gate_means=memory_gate_value or tmp_memory_gate_value)
        self._store_gate_means(out, info, d, memory_gate_value, tmp_memory_gate_value)
        return out.probs

    def mask_info(self, mask):
        """Host copy of a mask (active towers per level, edge matrices) cached per mask content
        so that the forward never synchronises on `bool(device_tensor)` like aread.py:272/297/309."""
        key = tuple((m.data_ptr(), m._version) if isinstance(m, torch.Tensor) else id(m) for m in mask)
        info = self._mask_cache.get(key)
        if info is None:
            if len(self._mask_cache) > 4096:
                self._mask_cache.clear()
            info = dense_ops.MaskInfo(hemp.to_numpy(mask), self.n_tower)
            info.keepalive = list(mask)     # pins the storage so a recycled data_ptr cannot alias the key
            self._mask_cache[key] = info
        return info

    def _store_gate_means(self, out, info, d, memory_gate_value, tmp_memory_gate_value):
        """aread.py:275-295: `tmp_tower_gate_values[l][t]` / `domain_tower_gate_values[d][l][t]` receive
        mean_b(softmax * mask column) for active towers and zeros for inactive ones."""
        if not (memory_gate_value or tmp_memory_gate_value):
            return
        for l in range(1, self.n_level):
            means = out.gate_means[l]                      # [n_tower[l-1], n_tower[l]], detached
            for t in range(self.n_tower[l]):
                v = means[:, t]
                if tmp_memory_gate_value:
                    self.tmp_tower_gate_values[l][t] = v.clone()
                if memory_gate_value:
                    self.domain_tower_gate_values[d][l][t].append(v.clone())

    def _record_unmasked_gates(self, x, gates, domain_i):
        """aread.py:187-200: per-domain batch means of the raw gate softmax."""
        if domain_i is not None:
            for l in range(1, self.n_level):
                means = gates[l].mean(dim=0)                # gates[l]: [B, n_tower[l-1], n_tower[l]]
                for t in range(self.n_tower[l]):
                    self.domain_tower_gate_values[domain_i][l][t].append(means[:, t].clone())
            return
        # mixed batch: per-domain means of every gate in ONE launch per level (csrc/mixed.cu domain_mean_kernel)
        # instead of n_domain x levels x towers boolean-index reductions; a domain without rows records NaN, the
        # mean of an empty selection, like the reference
        from . import embedding_ops, mixed_ops
        xi = embedding_ops.prepare_ids(x, self.embedding.embedding_dict.weight)
        for l in range(1, self.n_level):
            g = gates[l]                                                   # [B, n_{l-1}, n_l]
            n_prev, n_l = g.shape[1], g.shape[2]
            mean, count = mixed_ops.domain_means(g.reshape(g.shape[0], n_prev * n_l), xi, self.domain_idx, self.n_domain)
            mean = torch.where(count.view(-1, 1) > 0, mean, torch.full_like(mean, float('nan')))
            cols = mean.view(self.n_domain, n_prev, n_l).permute(0, 2, 1).contiguous()   # [d, t, n_prev]
            for d, per_tower in enumerate(cols.unbind(0)):
                log = self.domain_tower_gate_values[d][l]
                for t, v in enumerate(per_tower.unbind(0)):
                    log[t].append(v)

    # --------------------------------------------------------------------------- HEMP bookkeeping
    def add_eval_loss(self, loss_mean, d, mask_z):                                   # aread.py:324-328
        if len(self.eval_loss[d]) <= mask_z:
            self.eval_loss[d].append([loss_mean])
        else:
            self.eval_loss[d][mask_z].append(loss_mean)

    def update_all_mask(self, regroup_times=None, update_mode='best4single_domain'):   # aread.py:330-355
        print('\n============Update Mask============')
        if update_mode != 'best4single_domain':
            return
        n_cand = len(self.candidate_domain_mask[0])
        means, stds = [], []
        for d in range(self.n_domain):
            per_mask = [np.mean(self.eval_loss[d][z]) for z in range(n_cand)]
            self.domain_mask[d] = self.candidate_domain_mask[d][int(np.argmin(per_mask))]
            means.append(np.mean(per_mask))
            stds.append(np.std(per_mask))
        print('regroup_times: ', regroup_times,
              'current domain mask active ratio: ', self.count_current_active_ratio())
        print(f'loss_mean of different domain masks: {means}')
        print(f'loss_std of different domain masks: {stds}')
        users = [[] for _ in range(self.n_tower[1])]
        for d in range(self.n_domain):
            for t in np.nonzero(self.mask_info(self.domain_mask[d]).active[1])[0]:
                users[t].append(d)
        print(f'active domain num of each tower in the middle layer: {[len(u) for u in users]}')
        print('sample size training each tower in the middle layer: '
              f'{[sum(self.domain_size[u]) for u in users]}')
        print('============Finish Update Mask============')

    def prun_single_mask(self, d, current_mask, prun_ratio=0.05):                      # aread.py:357-381
        gate_values = [torch.stack(self.tmp_tower_gate_values[l], dim=1) for l in range(1, self.n_level)]
        threshold = 1
        for g in gate_values:
            live = g > 1e-8
            if live.any():
                threshold = min(threshold, torch.quantile(g[live].flatten(), prun_ratio))
        if threshold == 1:
            self.print_domain_mask(current_mask, all_edges=True)
            for i in range(self.n_level):
                print(f'level {i} gate_values: {self.tmp_tower_gate_values[i]}')
            raise ValueError('no valid tmp_tower_gate_values in candidate mask')
        before = copy.deepcopy(current_mask)
        for l in range(1, self.n_level):
            current_mask[l] = current_mask[l] & (gate_values[l - 1] >= threshold)
        valid = self.validate_mask(current_mask)
        self.tmp_tower_gate_values = [[None] * self.n_tower[l] for l in range(self.n_level)]
        return valid if valid[-1].any().item() else before

    def _empty_gate_log(self):
        return [[[] for _ in range(self.n_tower[l])] for l in range(self.n_level)] + \
               [[[] for _ in range(self.n_tower[-1])]]

    def reset_for_mask_update(self, d=None):                                          # aread.py:383-401
        if d is None:
            self.domain_tower_gate_values = [self._empty_gate_log() for _ in range(self.n_domain)]
            self.gate_value_threshold = [None] * self.n_domain
            self.candidate_domain_mask = [[] for _ in range(self.n_domain)]
            self.eval_loss = [[] for _ in range(self.n_domain)]
        else:
            self.domain_tower_gate_values[d] = self._empty_gate_log()
            self.gate_value_threshold[d] = None
            self.candidate_domain_mask[d] = []
            self.eval_loss[d] = []

    def mean_domain_tower_gate_values(self, d, get_threshold=None):                    # aread.py:403-430
        log = self.domain_tower_gate_values[d]
        if not isinstance(log[0], list):
            return
        dev = self.device
        mean_values = [torch.zeros(1, self.n_tower[0], dtype=torch.float32, device=dev)]
        for l in range(1, self.n_level):
            cols = []
            for t in range(self.n_tower[l]):
                if len(log[l][t]) == 0:
                    cols.append(torch.zeros(self.n_tower[l - 1], dtype=torch.float32, device=dev))
                else:
                    cols.append(torch.mean(torch.stack(log[l][t], dim=0), dim=0))
            mean_values.append(torch.stack(cols, dim=1))
        mean_values.append(torch.zeros(self.n_tower[-1], 1, dtype=torch.float32, device=dev))
        self.domain_tower_gate_values[d] = mean_values
        if get_threshold is not None:
            threshold = 1
            for ts in mean_values[1:-1]:
                live = ts > 1e-8
                if live.any():
                    threshold = min(threshold, torch.quantile(ts[live].flatten(), 1 - get_threshold))
            self.gate_value_threshold[d] = None if threshold == 1 else threshold

    def generate_mask(self, generate_mode='rand', d=None, init_active_percent=0.7, random_modify_sigma=0.2):
        """Candidate mask for one domain (aread.py:432-532).  Random draws are made in the
        reference's order so that a seeded schedule reproduces the same masks."""
        n_mat = self.n_level + 1
        if generate_mode == 'rand':
            while True:
                valid = hemp.validate_arrays(hemp.full_mask(self.n_tower, init_active_percent), tuple(self.n_tower))
                if valid[-1].any():
                    return hemp.as_device_mask(valid, self.device)
        if generate_mode == 'mask_norm_rand':
            original = hemp.to_numpy(self.domain_mask[d])
            n_active = sum(int(m.sum()) for m in original)
            while True:
                p = min(1, np.abs(np.random.normal(0, random_modify_sigma)))
                grow = n_active < self.edge_num * p
                cand = []
                for l in range(n_mat):
                    flip = np.random.rand(*original[l].shape) < p
                    cand.append(original[l] | flip if grow else original[l] ^ flip)
                valid = hemp.validate_arrays(cand, tuple(self.n_tower))
                changed = any(not np.array_equal(valid[l], original[l]) for l in range(n_mat))
                if changed and valid[-1].any():
                    return hemp.as_device_mask(valid, self.device)
        if generate_mode in ('max_gate', 'max_gate_norm_rand'):
            if not any(self.domain_tower_gate_values):
                raise ValueError('tower_gate_values is None')
            self.mean_domain_tower_gate_values(d, get_threshold=init_active_percent)
            if self.gate_value_threshold[d] is None:
                return self.generate_mask('rand', d, init_active_percent, random_modify_sigma)
            top = [t >= self.gate_value_threshold[d] for t in self.domain_tower_gate_values[d]]
            if generate_mode == 'max_gate':
                valid = self.validate_mask(top)
                if not valid[-1].any().item():
                    raise ValueError(f"mask generated for domain {d} in the 'max_gate' mode has no output")
                return valid
            p = min(1, np.abs(np.random.normal(0, random_modify_sigma)))
            while True:
                cand = [top[l] ^ (torch.rand(top[l].shape, device=self.device) < p) for l in range(n_mat)]
                valid = self.validate_mask(cand)
                if valid[-1].any().item():
                    return valid
        if generate_mode == 'mask_max_gate':
            if not any(self.domain_tower_gate_values):
                raise ValueError('tower_gate_values is None')
            self.mean_domain_tower_gate_values(d, get_threshold=init_active_percent)
            if self.gate_value_threshold[d] is None:
                top = self.generate_mask('rand', d, init_active_percent, random_modify_sigma)
            else:
                top = [t >= self.gate_value_threshold[d] for t in self.domain_tower_gate_values[d]]
            p = min(1, np.abs(np.random.normal(0, random_modify_sigma)))
            origin = self.domain_mask[d] if self.domain_mask[d] is not None else top
            shrink = (self.count_active_edge(d_mask=origin) * 1. / self.edge_num) > init_active_percent
            while True:
                cand = []
                for l in range(n_mat):
                    flip = torch.rand(top[l].shape, device=self.device) < p
                    merged = origin[l] | top[l]
                    cand.append(merged ^ flip if shrink else merged | flip)
                valid = self.validate_mask(cand)
                changed = any(not torch.all(valid[l] == origin[l]) for l in range(n_mat))
                if changed and valid[-1].any().item():
                    return valid
        raise ValueError(f"unknown generate_mode '{generate_mode}'")

    _ROLLBACK_PREFIXES = ('cn', 'cgc_layers', 'towers', 'tower_gates', 'towers_linear', 'output_layers',
                          'embedding', 'linear', 'reg_loss', 'regularization_weight')

    def save_model_state(self):                                                      # aread.py:534-543
        """Snapshot of every state entry under the roll-back prefixes.  The snapshot buffers persist from regroup to
        regroup (the reference deep-copies the whole table every time) and are refreshed with ONE multi-tensor
        copy launch; `model_state` stays the {key: tensor} dict the reference exposes."""
        pattern = re.compile('^(' + '|'.join(self._ROLLBACK_PREFIXES) + ')')
        live = {k: v for k, v in self.state_dict().items() if pattern.match(k)}
        fast = all(v.is_cuda and v.is_contiguous() and v.element_size() * v.numel() % 4 == 0 and v.numel() > 0
                   for v in live.values())
        if not fast:                                     # CPU module (host-logic tests): plain clones
            self.model_state = {k: v.detach().clone() for k, v in live.items()}
            object.__setattr__(self, "_rollback", None)
            return
        snap = self.model_state
        if (snap is None or snap.keys() != live.keys() or
                any(snap[k].shape != v.shape or snap[k].dtype != v.dtype or snap[k].device != v.device
                    for k, v in live.items())):
            snap = self.model_state = {k: torch.empty_like(v) for k, v in live.items()}
            object.__setattr__(self, "_rollback", None)
        rb = getattr(self, "_rollback", None)
        keys = list(live)
        ptrs = [live[k].data_ptr() for k in keys]
        if rb is None or rb["ptrs"] != ptrs or any(rb["snap"][i] is not snap[k] for i, k in enumerate(keys)):
            from .optim import MultiCopy
            dsts, srcs = [snap[k] for k in keys], [live[k].detach() for k in keys]
            owners = []
            for k in keys:                               # (module, attribute) of every entry: looked up again at
                path, _, leaf = k.rpartition('.')        # each restore, so a replaced Parameter is noticed
                owners.append((self.get_submodule(path) if path else self, leaf))
            rb = {"save": MultiCopy(dsts, srcs), "load": MultiCopy(srcs, dsts), "ptrs": ptrs, "snap": dsts,
                  "owners": owners}
            object.__setattr__(self, "_rollback", rb)
        rb["save"].run()

    def load_model_state(self):                                                      # aread.py:545-546
        """`load_state_dict(model_state, strict=False)` of the reference = in-place copies into the live tensors;
        here as one launch over all of them (falls back to load_state_dict when storage moved since the save)."""
        rb = getattr(self, "_rollback", None)
        if rb is not None and all(getattr(m, leaf).data_ptr() == p for (m, leaf), p in zip(rb["owners"], rb["ptrs"])):
            rb["load"].run()
        else:
            self.load_state_dict(self.model_state, strict=False)

    def create_single_full_mask(self, fill_value=0):                                 # aread.py:548-568
        return hemp.full_mask(self.n_tower, fill_value)

    def validate_mask(self, mask, add_input=True, add_output=True, remove_hidden=True):   # aread.py:570-605
        return hemp.validate(mask, self.n_tower, add_input, add_output, remove_hidden)

    def create_domain_mask(self, cluster_z):                                         # aread.py:607-638
        """Initial masks from a hierarchical-clustering linkage matrix (unused by run.py, kept)."""
        masks = [hemp.full_mask(self.n_tower, 0) for _ in range(self.n_domain)]
        members = [[i] for i in range(self.n_domain)]
        alive = list(range(self.n_domain))
        level_clusters = [None] * self.n_level
        for i in range(self.n_domain - self.n_tower[0]):
            a, b = int(cluster_z[i][0]), int(cluster_z[i][1])
            members.append(members[a] + members[b])
            alive.append(i + self.n_domain)
            alive.remove(a)
            alive.remove(b)
            if len(alive) in self.n_tower:
                level_clusters[self.n_tower.index(len(alive))] = list(alive)
        for l in range(self.n_level):
            for t in range(self.n_tower[l]):
                cluster = members[level_clusters[l][t]]
                self.tower2cluster[l][t] = cluster
                for d in cluster:
                    masks[d][l + 1][t, :] = True
        self.domain_mask = [hemp.as_device_mask(hemp.validate_arrays(m, tuple(self.n_tower)), self.device)
                            for m in masks]

    def print_domain_mask(self, d_mask=None, d=None, all_edges=False):                 # aread.py:640-662
        mask = hemp.to_numpy(d_mask if d_mask is not None else self.domain_mask[d])
        print('level 0 towers:', np.nonzero(mask[0])[1])
        if all_edges:
            for l in range(1, self.n_level):
                print(f'========= level {l} =========')
                for t in range(self.n_tower[l]):
                    feeders = np.nonzero(mask[l][:, t])[0]
                    print(f'last level input towers of tower {t}:', feeders if len(feeders) else None)
            print('========= level finish =========')
        else:
            for l in range(1, self.n_level):
                print(f'level {l} used last level towers:', np.nonzero(np.any(mask[l], axis=1))[0])
        print('the last level output towers:', np.nonzero(mask[-1])[0])

    def count_current_active_ratio(self):                                            # aread.py:664-669
        return sum(self.count_active_edge(d=d) * 1. / self.edge_num for d in range(self.n_domain)) / self.n_domain

    def count_active_edge(self, d=None, d_mask=None):                                  # aread.py:671-680
        return hemp.count_edges(d_mask if d_mask is not None else self.domain_mask[d])
