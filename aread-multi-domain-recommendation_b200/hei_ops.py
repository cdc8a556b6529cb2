"""Torch-tensor wrappers over the fused HEI tower-layer entry points (csrc/hei.cu); no arithmetic here.

`saved` is the [4, width] fp32 block (mean, rstd, scale, shift) the BatchNorm kernels share; `src_saved`
is the same block of the layer below when the layer input is that layer's pre-activation."""
import ctypes

import torch

from . import _lib
from . import _mem
from . import dense_kernels as dk
from .dense_kernels import BN_EPS, BN_MOMENTUM, _bn_workspace, _ptr, _rows


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def supported(groups, k, n):
    return bool(_lib.load().aread_hei_layer_supported(groups, k, n))


def set_path(tensor_cores_fwd=-1, tensor_cores_bwd=-1):
    """Which implementation the tower layers run from now on: 1 = tcgen05 where the shape packs, 0 = CUDA cores,
    -1 = the environment's default (AREAD_HEI_TC / AREAD_HEI_TC_BWD).  Recorded CUDA graphs keep what they captured."""
    _lib.load().aread_hei_set_path(int(tensor_cores_fwd), int(tensor_cores_bwd))


def path(m, groups, k, n):
    """(forward on tensor cores, backward on tensor cores) for a call of this shape."""
    bits = int(_lib.load().aread_hei_layer_path(int(m), groups, k, n))
    return bool(bits & 1), bool(bits & 2)


def _workspace(device, m, groups, k, n):
    need = int(_lib.load().aread_hei_layer_workspace_bytes(m, groups, k, n))
    return _mem.workspace("hei", device, need)


def layer_fwd(src, src_saved, src_salt, weight, bias, gamma, beta, running_mean, running_var, groups, k, n, training,
              bn_skip, p, seed):
    """src [m, groups * k] -> (z [m, groups * n], saved [4, groups * n])."""
    m = src.shape[0]
    dev = src.device
    z = _mem.empty((m, groups * n), torch.float32, dev)
    saved = _mem.empty((4, groups * n), torch.float32, dev)
    ws = _workspace(dev, m, groups, k, n)
    sp = _rows(src_saved, 4) if src_saved is not None else (None,) * 4
    args = _lib.HeiLayerFwdArgs(
        m, groups, k, n, 1 if training else 0, 1 if bn_skip else 0, BN_MOMENTUM, BN_EPS, src.data_ptr(), src.stride(0),
        sp[2], sp[3], float(p) if training else 0.0, src_salt, seed, weight.data_ptr(), _ptr(bias), _ptr(gamma),
        _ptr(beta), _ptr(running_mean), _ptr(running_var), z.data_ptr(), *_rows(saved, 4), ws.data_ptr(), ws.numel(),
        dk.SEED_PTR)
    _lib.check(_lib.load().aread_hei_layer_fwd(ctypes.byref(args), _stream(dev)))
    return z, saved


def bn_apply(z, saved, training, p, seed, salt):
    """dropout(relu(z * scale + shift)) as fp32 [m, width]."""
    m, width = z.shape
    out = _mem.empty((m, width), torch.float32, z.device)
    args = _lib.BnActArgs(m, width, 1 if training else 0, 0, BN_MOMENTUM, BN_EPS, float(p) if training else 0.0, seed,
                          salt, z.data_ptr(), z.stride(0), None, None, None, None, *_rows(saved, 4), out.data_ptr(),
                          None, width, None, 0, None, dk.SEED_PTR)
    _lib.check(_lib.load().aread_bn_act_apply(ctypes.byref(args), _stream(z.device)))
    return out


def bn_bwd_coef(z, d_out, saved, bn_skip, p, seed, salt):
    """(coef [2, width], grads [3, width] = d_gamma, d_beta, d_bias) of out = dropout(relu(bn(z)))."""
    m, width = z.shape
    coef = _mem.empty((2, width), torch.float32, z.device)
    grads = torch.empty((3, width), dtype=torch.float32, device=z.device)        # parameter gradients: never arena
    ws = _bn_workspace(z.device, width)
    args = _lib.BnActBwdArgs(m, width, 1 if bn_skip else 0, float(p), salt, seed, z.data_ptr(), z.stride(0),
                             d_out.data_ptr(), d_out.stride(0), *_rows(saved, 4), *_rows(grads, 3), None, None, width,
                             ws.data_ptr(), ws.numel(), None, dk.SEED_PTR)
    _lib.check(_lib.load().aread_bn_bwd_coef(ctypes.byref(args), ctypes.c_void_p(coef.data_ptr()), _stream(z.device)))
    return coef, grads


def layer_bwd(z, d_out, saved, coef, p, salt, seed, bn_skip, src, src_saved, src_salt, weight, groups, k, n):
    """-> (d_in [m, groups * k], d_w [groups, n, k], src_coef or None, src_grads [3, groups * k] or None)."""
    m = z.shape[0]
    dev = z.device
    d_in = _mem.empty((m, groups * k), torch.float32, dev)
    d_w = torch.empty((groups, n, k), dtype=torch.float32, device=dev)           # parameter gradient: never arena
    has = src_saved is not None
    src_coef = _mem.empty((2, groups * k), torch.float32, dev) if has else None
    src_grads = torch.empty((3, groups * k), dtype=torch.float32, device=dev) if has else None
    ws = _workspace(dev, m, groups, k, n)
    sp = _rows(src_saved, 4) if has else (None,) * 4
    gp = _rows(src_grads, 3) if has else (None,) * 3
    args = _lib.HeiLayerBwdArgs(
        m, groups, k, n, 1 if bn_skip else 0, float(p), salt, seed, z.data_ptr(), d_out.data_ptr(), *_rows(saved, 4),
        coef.data_ptr(), src.data_ptr(), src.stride(0), sp[2], sp[3], sp[0], sp[1], float(p), src_salt,
        weight.data_ptr(), d_in.data_ptr(), d_w.data_ptr(), _ptr(src_coef), *gp, ws.data_ptr(), ws.numel(),
        dk.SEED_PTR)
    _lib.check(_lib.load().aread_hei_layer_bwd(ctypes.byref(args), _stream(dev)))
    return d_in, d_w, src_coef, src_grads
