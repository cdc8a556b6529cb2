"""Torch-tensor wrappers over aread_tower_linear / aread_tower_wgrad (csrc/tower.cu): grouped small Linear layers
in fp32 for the gate logits, the wider-than-64 tower layers and a few one-off products of the fused node."""
import ctypes

import torch

from . import _lib
from . import _mem


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
    ws = _mem.workspace("tower_wgrad", dz.device, need)
    args = _lib.TowerWgradArgs(B, G, N, K, dz.data_ptr(), G * N, x.data_ptr(), ld, gs, d_w.data_ptr(), ws.data_ptr(),
                               ws.numel())
    _lib.check(_lib.load().aread_tower_wgrad(ctypes.byref(args), _stream(dz.device)))
    return d_w
