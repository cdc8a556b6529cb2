"""Grouped tower Linear kernels (C ABI: aread_tower_linear / aread_tower_wgrad) and the L2 regulariser
kernel against plain torch fp32 references.  Tolerance: fp32 summation-order round-off."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
to = importlib.import_module("aread-multi-domain-recommendation_b200.tower_ops")
ro = importlib.import_module("aread-multi-domain-recommendation_b200.reg_ops")
DEV = "cuda:0"

SHAPES = [(1000, 3, 64, 64), (1000, 3, 64, 32), (777, 6, 32, 16), (129, 12, 16, 8), (37, 4, 8, 4), (1, 2, 8, 8),
          (5000, 12, 16, 16), (300, 5, 20, 12), (64, 1, 128, 3)]


@pytest.mark.parametrize("m,g,k,n", SHAPES)
def test_tower_linear_forward_and_gradients(m, g, k, n):
    gen = torch.Generator(device=DEV).manual_seed(m + k + n)
    x = torch.randn(m, g, k, device=DEV, generator=gen)
    w = torch.randn(g, n, k, device=DEV, generator=gen) / k ** 0.5
    b = torch.randn(g, n, device=DEV, generator=gen)
    z = to.tower_linear(x, w, b, n)
    ref = torch.einsum("bgk,gnk->bgn", x, w) + b
    torch.testing.assert_close(z, ref, rtol=1e-5, atol=1e-5)
    dz = torch.randn(m, g, n, device=DEV, generator=gen)
    d_x = to.tower_linear(dz, w, None, k, weight_is_out_by_in=False)
    torch.testing.assert_close(d_x, torch.einsum("bgn,gnk->bgk", dz, w), rtol=1e-5, atol=1e-5)
    d_w = to.tower_wgrad(dz, x)
    ref_w = torch.einsum("bgn,bgk->gnk", dz.double(), x.double()).float()
    torch.testing.assert_close(d_w, ref_w, rtol=1e-4, atol=1e-4 * m ** 0.5)
    assert torch.equal(d_w, to.tower_wgrad(dz, x))                      # fixed reduction order


def test_l2_regulariser_kernel():
    gen = torch.Generator(device=DEV).manual_seed(1)
    shapes = [(100000, 32), (1, 288), (256, 288), (256,), (7,), (4097,), (64, 64)]
    l2s = [1e-5, 1e-5, 2e-5, 1e-5, 3e-5, 1e-5, 1e-5]
    ws = [torch.randn(*s, device=DEV, generator=gen).requires_grad_(True) for s in shapes]
    reg = [([ws[0]], 0.0, l2s[0]), ([("fc.weight", ws[1])], 0.0, l2s[1])] + \
          [([w], 0.0, l2) for w, l2 in zip(ws[2:], l2s[2:])]
    out = ro.regularization_loss(reg, torch.device(DEV))
    ref = sum(torch.sum(l2 * torch.square(w.double())) for w, l2 in zip(ws, l2s))
    assert tuple(out.shape) == (1,)
    torch.testing.assert_close(out.double(), ref.reshape(1), rtol=1e-6, atol=0)
    (out * 3.0).backward()
    for w, l2 in zip(ws, l2s):
        torch.testing.assert_close(w.grad, 3.0 * 2 * l2 * w.detach(), rtol=1e-6, atol=0)
    again = ro.regularization_loss(reg, torch.device(DEV))
    assert torch.equal(out, again)                                       # bit-reproducible
    # the backward's buffer is reused from step to step: .grad must never alias it, and two live graphs of the
    # same tensor list must not share it
    first = [w.grad.clone() for w in ws]
    a, b = ro.regularization_loss(reg, torch.device(DEV)), ro.regularization_loss(reg, torch.device(DEV))
    for w in ws:
        w.grad = None
    (0.5 * a + 2.0 * b).backward()
    for w, l2, g3 in zip(ws, l2s, first):
        torch.testing.assert_close(w.grad, 2.5 * 2 * l2 * w.detach(), rtol=1e-6, atol=0)
        torch.testing.assert_close(g3, 3.0 * 2 * l2 * w.detach(), rtol=1e-6, atol=0)
    kept = [w.grad for w in ws]
    snap = [g.clone() for g in kept]
    ro.regularization_loss(reg, torch.device(DEV)).backward()            # accumulates into .grad, does not replace it
    for w, g, s0, l2 in zip(ws, kept, snap, l2s):
        assert w.grad is g
        torch.testing.assert_close(g, s0 + 2 * l2 * w.detach(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("T,m", [(1, 1), (3, 37), (9, 4096), (12, 70001)])
def test_bagging_bce_matches_bceloss(T, m):
    lo = importlib.import_module("aread-multi-domain-recommendation_b200.loss_ops")
    gen = torch.Generator(device=DEV).manual_seed(T * 1000 + m)
    probs = torch.sigmoid(4 * torch.randn(T, m, device=DEV, generator=gen))
    probs[0, 0] = 0.0                                  # log clamped at -100, like BCELoss
    if m > 1:
        probs[-1, 1] = 1.0
    y = (torch.rand(m, 1, device=DEV, generator=gen) < 0.3).to(torch.int16)
    p_ref = probs.clone().requires_grad_(True)
    crit = torch.nn.BCELoss()
    ref = sum(crit(p, y.squeeze(1).float()) for p in p_ref.unbind(0)) / T
    (ref * 1.5).backward()
    p_mine = probs.clone().requires_grad_(True)
    mine = lo.bagging_bce(p_mine, y)
    (mine * 1.5).backward()
    assert mine.shape == ref.shape
    torch.testing.assert_close(mine, ref, rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(p_mine.grad, p_ref.grad, rtol=1e-6, atol=0)
    assert torch.equal(lo.bagging_bce(probs, y), mine.detach())          # bit-reproducible


@pytest.mark.parametrize("m,T,W", [(1, 1, 8), (37, 3, 8), (5000, 9, 8), (70001, 12, 4)])
def test_head_kernels_match_autograd(m, T, W):
    import ctypes
    _lib = importlib.import_module("aread-multi-domain-recommendation_b200._lib")
    gen = torch.Generator(device=DEV).manual_seed(m + T)
    head_cross = torch.randn(m, T, device=DEV, generator=gen)
    lin = torch.randn(m, device=DEV, generator=gen)
    h = torch.randn(m, T, W, device=DEV, generator=gen).requires_grad_(True)
    w = torch.randn(T, W, device=DEV, generator=gen).requires_grad_(True)
    hc = head_cross.clone().requires_grad_(True)
    ln = lin.clone().requires_grad_(True)
    ref = torch.sigmoid(hc + (h * w).sum(dim=2) + ln.unsqueeze(1)).t().contiguous()
    d_probs = torch.randn(T, m, device=DEV, generator=gen)
    ref.backward(d_probs)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    probs = torch.empty(T, m, device=DEV)
    a = _lib.HeadArgs(m, T, W, head_cross.data_ptr(), lin.data_ptr(), h.data_ptr(), w.data_ptr(), probs.data_ptr(),
                      None, None, None, None, None, None, 0)
    _lib.check(_lib.load().aread_head(ctypes.byref(a), stream))
    torch.testing.assert_close(probs, ref.detach(), rtol=1e-5, atol=1e-6)
    dz, d_lin, d_h, d_w = (torch.empty(m, T, device=DEV), torch.empty(m, device=DEV), torch.empty(m, T, W, device=DEV),
                           torch.empty(T, W, device=DEV))
    ws = torch.empty(int(_lib.load().aread_head_workspace_bytes(T, W)), dtype=torch.uint8, device=DEV)
    a = _lib.HeadArgs(m, T, W, None, None, h.data_ptr(), w.data_ptr(), probs.data_ptr(), d_probs.data_ptr(),
                      dz.data_ptr(), d_lin.data_ptr(), d_h.data_ptr(), d_w.data_ptr(), ws.data_ptr(), ws.numel())
    _lib.check(_lib.load().aread_head(ctypes.byref(a), stream))
    torch.testing.assert_close(dz, hc.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_lin, ln.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d_h, h.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_w, w.grad, rtol=1e-4, atol=1e-4 * max(1.0, m ** 0.5))
    first = d_w.clone()
    _lib.check(_lib.load().aread_head(ctypes.byref(a), stream))
    assert torch.equal(d_w, first)                                       # fixed reduction order
