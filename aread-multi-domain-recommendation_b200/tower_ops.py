"""One HEI tower layer (Linear -> BatchNorm1d -> ReLU -> Dropout) for all towers of a level that run
under the current HEMP mask, as one autograd node on csrc/tower.cu + csrc/bn_act.cu.

Activations are compact: [B, n_active, width] holds only the towers that run, in ascending tower
order; towers pruned by the mask cost nothing and their parameters receive no gradient (None), like
the reference (model/aread.py:272-281, 297-300, 319-321)."""
import ctypes

import torch

from . import _lib
from . import _mem
from . import dense_kernels as dk

_WS = {}


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def tower_linear(x, weight, bias, out_width, weight_is_out_by_in=True, groups=None):
    """x [B, G, I] fp32 (or [B, I] shared by `groups` groups), weight [G, ...] contiguous ->
    [B, G, out_width] fp32."""
    if x.dim() == 2:
        B, I = x.shape
        G, ld, gs = groups, I, 0
    else:
        B, G, I = x.shape
        ld, gs = G * I, I
    out = _mem.empty((B, G, out_width), torch.float32, x.device)
    args = _lib.TowerLinearArgs(B, G, I, out_width, 1 if weight_is_out_by_in else 0, x.data_ptr(), ld, gs,
                                weight.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                G * out_width)
    _lib.check(_lib.load().aread_tower_linear(ctypes.byref(args), _stream(x.device)))
    return out


def tower_wgrad(dz, x):
    """dz [B, G, N], x [B, G, K] (or [B, K] shared by all groups) -> d_w [G, N, K] fp32."""
    B, G, N = dz.shape
    K = x.shape[-1]
    ld, gs = (K, 0) if x.dim() == 2 else (G * K, K)
    d_w = torch.empty((G, N, K), dtype=torch.float32, device=dz.device)
    need = int(_lib.load().aread_tower_wgrad_workspace_bytes(B, G, N, K))
    ws = _WS.get(dz.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dz.device)
        _WS[dz.device] = ws
    args = _lib.TowerWgradArgs(B, G, N, K, dz.data_ptr(), G * N, x.data_ptr(), ld, gs, d_w.data_ptr(), ws.data_ptr(),
                               ws.numel())
    _lib.check(_lib.load().aread_tower_wgrad(ctypes.byref(args), _stream(dz.device)))
    return d_w


class TowerLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg, *params):
        L, active, training, p, seed = cfg["layer"], cfg["active"], cfg["training"], cfg["dropout"], cfg["seed"]
        B, na, K = x.shape
        G, N = L.groups, L.n
        full = na == G
        x = x.contiguous()
        if full:
            w, bias, gamma, beta = L.weight.flat, L.bias.flat, L.gamma.flat, L.beta.flat
            rm, rv = L.running_mean.flat, L.running_var.flat
        else:
            idx = cfg["index"]
            w, bias, gamma, beta = (t.flat.index_select(0, idx) for t in (L.weight, L.bias, L.gamma, L.beta))
            rm, rv = L.running_mean.flat.index_select(0, idx), L.running_var.flat.index_select(0, idx)
        z = tower_linear(x, w, bias, N)
        bn_skip = B == 1
        out, stats = dk.bn_act_fwd(z.view(B, na * N), gamma.reshape(-1), beta.reshape(-1), rm.view(-1), rv.view(-1),
                                   training, bn_skip, p, seed, L.salt, torch.float32)
        if training and not bn_skip:
            if not full:
                L.running_mean.flat.index_copy_(0, idx, rm)
                L.running_var.flat.index_copy_(0, idx, rv)
            tracked = L.tracked
            torch._foreach_add_([tracked[t] for t in active], 1)
        ctx.cfg, ctx.bn_skip = cfg, bn_skip
        ctx.save_for_backward(x, z, stats, w)
        return out.view(B, na, N)

    @staticmethod
    def backward(ctx, d_out):
        cfg = ctx.cfg
        L, active, training, seed = cfg["layer"], cfg["active"], cfg["training"], cfg["seed"]
        p = cfg["dropout"] if training else 0.0
        x, z, stats, w = ctx.saved_tensors
        B, na, K = x.shape
        G, N = L.groups, L.n
        dz, d_gamma, d_beta, d_bias = dk.bn_act_bwd(z.view(B, na * N), d_out.contiguous().view(B, na * N), stats,
                                                    ctx.bn_skip, p, seed, L.salt, torch.float32)
        dz = dz.view(B, na, N)
        d_w = tower_wgrad(dz, x)
        d_x = tower_linear(dz, w, None, K, weight_is_out_by_in=False) if ctx.needs_input_grad[0] else None
        per_tower = [None] * (4 * G)          # order of ExpertLayer.params: weights, biases, gammas, betas
        d_bias, d_gamma, d_beta = d_bias.view(na, N), d_gamma.view(na, N), d_beta.view(na, N)
        for i, t in enumerate(active):
            per_tower[t] = d_w[i]
            per_tower[G + t] = d_bias[i]
            per_tower[2 * G + t] = d_gamma[i]
            per_tower[3 * G + t] = d_beta[i]
        return (d_x, None, *per_tower)


def tower_layer(x, layer, active, index, training, dropout, seed):
    """x: [B, n_active, K] -> [B, n_active, N] for the towers listed in `active` (ascending)."""
    cfg = {"layer": layer, "active": active, "index": index, "training": training, "dropout": dropout, "seed": seed}
    return TowerLayerFn.apply(x, cfg, *layer.params)
