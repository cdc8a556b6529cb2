"""Fused HEI tower-layer kernels (C ABI: aread_hei_layer_fwd / _bwd, aread_bn_act_apply, aread_bn_bwd_coef) against
the same two-layer tower stack written with torch ops and differentiated by autograd (fp32; the dropout masks come
from the library's own counter stream through aread_dropout_mask).  Tolerance: fp32 summation-order noise for the
CUDA-core kernels (csrc/hei.cu); the tensor-core kernels (csrc/hei_tc.cu, rows >= 512 and block-packable widths)
split every operand into bf16 hi + lo and drop lo.lo, |dz| <= ~1e-4 at these magnitudes.

ReLU boundary: an element whose BatchNorm output is within the forward tolerance of zero may be rectified
differently by two correct implementations, and that flips a whole gradient term.  The reference therefore takes
the on/off decision from the kernel's own pre-activation -- after checking that the two only disagree where the
reference value is within 1e-3 of zero -- so the gradients are compared under one activation pattern."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ho = importlib.import_module("aread-multi-domain-recommendation_b200.hei_ops")
dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")


def _ref_layer(x, w, b, gamma, beta, mask, p, bn_skip, kernel_y=None):
    """x [m, G, K] -> (z [m, G, N], act [m, G, N]) with train-mode BatchNorm over the batch.  `kernel_y`: the
    kernel's BatchNorm output, whose sign decides the ReLU (see the module docstring)."""
    z = torch.einsum("bgk,gnk->bgn", x, w) + b
    if bn_skip:
        y = z
    else:
        mean, var = z.mean(dim=0), z.var(dim=0, unbiased=False)
        y = (z - mean) / torch.sqrt(var + 1e-5) * gamma + beta
    if kernel_y is None:
        a = torch.relu(y)
    else:
        on = kernel_y > 0
        differ = on != (y.detach() > 0)
        assert float(y.detach().abs()[differ].max() if differ.any() else 0.0) < 1e-3, "ReLU decision off the boundary"
        a = y * on
    if p > 0:
        a = a * mask / (1 - p)
    return z, a


@pytest.fixture(autouse=True)
def _default_path():
    yield
    ho.set_path(-1, -1)


# tensor cores: rows >= 512 and (k, n) that pack into 64-column blocks; several units per CTA from ~20,000 rows
# (towers x rows: one block 148 CTAs, three blocks 49 each), partial last blocks (G not a multiple of 64 / k)
CASES = [(1000, 3, (64, 64, 32), 0.0), (777, 5, (32, 32, 16), 0.2), (4099, 9, (16, 16, 8), 0.2),
         (37, 2, (64, 64, 32), 0.2), (70000, 2, (16, 16, 8), 0.0), (2, 1, (8, 12, 4), 0.0),
         (513, 4, (20, 10, 6), 0.2), (1, 3, (16, 16, 8), 0.0), (40000, 3, (64, 64, 32), 0.2),
         (70000, 4, (16, 16, 8), 0.2), (30000, 7, (32, 32, 16), 0.2)]


def test_paths_are_the_documented_ones():
    assert ho.path(65536, 3, 64, 64) == (True, True) and ho.path(65536, 12, 16, 8) == (True, True)
    assert ho.path(300, 3, 64, 64) == (False, False), "under 512 rows: CUDA cores"
    assert ho.path(65536, 4, 20, 10) == (False, False), "widths that do not pack into 64-column blocks"
    ho.set_path(0, 0)
    assert ho.path(65536, 3, 64, 64) == (False, False)
    ho.set_path(1, 0)
    assert ho.path(65536, 3, 64, 64) == (True, False)


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("m,G,dims,p", CASES)
def test_two_layer_stack_matches_autograd(m, G, dims, p, tensor_cores):
    ho.set_path(1 if tensor_cores else 0, 1 if tensor_cores else 0)
    if tensor_cores and ho.path(m, G, dims[0], dims[1]) == (False, False) and ho.path(m, G, dims[1], dims[2]) == (False, False):
        pytest.skip("shape runs on the CUDA cores either way")
    K, N1, N2 = dims
    gen = torch.Generator(device=DEV).manual_seed(m + G)

    def rnd(*s):
        return torch.randn(*s, device=DEV, generator=gen)

    x = rnd(m, G, K).requires_grad_(True)
    w1, b1 = (0.3 * rnd(G, N1, K)).requires_grad_(True), rnd(G, N1).requires_grad_(True)
    g1, be1 = (1 + 0.1 * rnd(G, N1)).requires_grad_(True), (0.1 * rnd(G, N1)).requires_grad_(True)
    w2, b2 = (0.3 * rnd(G, N2, N1)).requires_grad_(True), rnd(G, N2).requires_grad_(True)
    g2, be2 = (1 + 0.1 * rnd(G, N2)).requires_grad_(True), (0.1 * rnd(G, N2)).requires_grad_(True)
    d_u = rnd(m, G, N2)
    seed, salt1, salt2 = 1234567, 0x2001, 0x2002
    bn_skip = m == 1
    rm1, rv1 = torch.zeros(G * N1, device=DEV), torch.ones(G * N1, device=DEV)
    rm2, rv2 = torch.zeros(G * N2, device=DEV), torch.ones(G * N2, device=DEV)
    mask1 = dk.dropout_mask(seed, salt1, (m, G, N1), p, DEV).float()
    mask2 = dk.dropout_mask(seed, salt2, (m, G, N2), p, DEV).float()

    with torch.no_grad():
        src = x.detach().reshape(m, G * K)
        z1, s1 = ho.layer_fwd(src, None, 0, w1.detach(), b1.detach(), g1.detach().reshape(-1), be1.detach().reshape(-1),
                              rm1, rv1, G, K, N1, True, bn_skip, p, seed)
        z2, s2 = ho.layer_fwd(z1, s1, salt1, w2.detach(), b2.detach(), g2.detach().reshape(-1),
                              be2.detach().reshape(-1), rm2, rv2, G, N1, N2, True, bn_skip, p, seed)
        u = ho.bn_apply(z2, s2, True, p, seed, salt2)
        # saved = (mean, rstd, scale, shift); addcmul is the kernels' fused multiply-add
        y1_k = torch.addcmul(s1[3].expand_as(z1), z1, s1[2].expand_as(z1)).view(m, G, N1)
        y2_k = torch.addcmul(s2[3].expand_as(z2), z2, s2[2].expand_as(z2)).view(m, G, N2)

    z1_ref, a1_ref = _ref_layer(x, w1, b1, g1, be1, mask1, p, bn_skip, y1_k)
    z2_ref, u_ref = _ref_layer(a1_ref, w2, b2, g2, be2, mask2, p, bn_skip, y2_k)
    (u_ref * d_u).sum().backward()

    with torch.no_grad():
        tol = dict(rtol=2e-4, atol=2e-4)
        torch.testing.assert_close(z1.view(m, G, N1), z1_ref, **tol)
        torch.testing.assert_close(z2.view(m, G, N2), z2_ref, **tol)
        torch.testing.assert_close(u.view(m, G, N2), u_ref, **tol)
        if not bn_skip:
            torch.testing.assert_close(s1[0].view(G, N1), z1_ref.mean(dim=0), **tol)
            torch.testing.assert_close(rv2.view(G, N2), 0.9 + 0.1 * z2_ref.var(dim=0, unbiased=True), **tol)
            torch.testing.assert_close(rm2.view(G, N2), 0.1 * z2_ref.mean(dim=0), **tol)

        d_out = d_u.reshape(m, G * N2).contiguous()
        coef, g3 = ho.bn_bwd_coef(z2, d_out, s2, bn_skip, p, seed, salt2)
        d_a1, d_w2, coef1, g3_1 = ho.layer_bwd(z2, d_out, s2, coef, p, salt2, seed, bn_skip, z1, s1, salt1,
                                               w2.detach(), G, N1, N2)
        d_x, d_w1, none_c, none_g = ho.layer_bwd(z1, d_a1, s1, coef1, p, salt1, seed, bn_skip, src, None, 0,
                                                 w1.detach(), G, K, N1)
        assert none_c is None and none_g is None
        gt = dict(rtol=1e-3, atol=2e-4 * max(1.0, m ** 0.5))
        torch.testing.assert_close(d_w2, w2.grad, **gt)
        torch.testing.assert_close(d_w1, w1.grad, **gt)
        torch.testing.assert_close(d_x.view(m, G, K), x.grad, rtol=1e-3, atol=2e-4)
        if bn_skip:
            torch.testing.assert_close(g3[2].view(G, N2), b2.grad, **gt)
            torch.testing.assert_close(g3_1[2].view(G, N1), b1.grad, **gt)
        else:
            torch.testing.assert_close(g3[0].view(G, N2), g2.grad, **gt)
            torch.testing.assert_close(g3[1].view(G, N2), be2.grad, **gt)
            torch.testing.assert_close(g3_1[0].view(G, N1), g1.grad, **gt)
            torch.testing.assert_close(g3_1[1].view(G, N1), be1.grad, **gt)
        # fixed reduction order: bit-identical on a second run
        again = ho.layer_bwd(z2, d_out, s2, coef, p, salt2, seed, bn_skip, z1, s1, salt1, w2.detach(), G, N1, N2)
        assert torch.equal(again[1], d_w2) and torch.equal(again[0], d_a1) and torch.equal(again[2], coef1)


@pytest.mark.parametrize("m", [300, 3000])
def test_eval_mode_uses_running_statistics(m):
    G, K, N = 4, 32, 16
    gen = torch.Generator(device=DEV).manual_seed(5)

    def rnd(*s):
        return torch.randn(*s, device=DEV, generator=gen)

    x, w, b = rnd(m, G * K), 0.3 * rnd(G, N, K), rnd(G, N)
    gamma, beta = 1 + 0.1 * rnd(G * N), 0.1 * rnd(G * N)
    rm, rv = 0.2 * rnd(G * N), 0.5 + torch.rand(G * N, device=DEV, generator=gen)
    rm0, rv0 = rm.clone(), rv.clone()
    z, s = ho.layer_fwd(x, None, 0, w, b, gamma, beta, rm, rv, G, K, N, False, False, 0.3, 99)
    out = ho.bn_apply(z, s, False, 0.3, 99, 7)
    z_ref = torch.einsum("bgk,gnk->bgn", x.view(m, G, K), w) + b
    ref = torch.relu(F.batch_norm(z_ref.reshape(m, G * N), rm0, rv0, gamma, beta, False, 0.1, 1e-5))
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)
    assert torch.equal(rm, rm0) and torch.equal(rv, rv0)
