"""BatchNorm/ReLU/Dropout and MMoE-mixture kernels (C ABI) against a plain torch fp32 reference of the
same ops.  Tolerance: fp32 round-off (column sums are accumulated in a different order)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")
DEV = "cuda:0"
SEED, SALT = 0x1234567890ABCDEF, 77


def make(m, width, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    z = torch.randn(m, width, device=DEV, generator=g) * 1.7 + 0.3
    gamma = 1 + 0.1 * torch.randn(width, device=DEV, generator=g)
    beta = 0.1 * torch.randn(width, device=DEV, generator=g)
    rm = 0.1 * torch.randn(width, device=DEV, generator=g)
    rv = 0.5 + torch.rand(width, device=DEV, generator=g)
    return z, gamma, beta, rm, rv


@pytest.mark.parametrize("m", [2, 37, 5000])
@pytest.mark.parametrize("width", [1024, 512, 96, 24, 8, 3])
def test_bn_act_forward_training(m, width):
    z, gamma, beta, rm, rv = make(m, width, m + width)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.relu(F.batch_norm(z, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5))
    out, saved = dk.bn_act_fwd(z, gamma, beta, rm, rv, True, False, 0.0, SEED, SALT, torch.float32)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(rm, rm_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv, rv_ref, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(saved[0], z.mean(0), rtol=1e-5, atol=1e-5)
    out16, _ = dk.bn_act_fwd(z, gamma, beta, rm.clone(), rv.clone(), True, False, 0.0, SEED, SALT, torch.bfloat16)
    assert torch.equal(out16, out.to(torch.bfloat16))                    # same statistics, one extra rounding


def test_bn_act_forward_eval_and_skip():
    z, gamma, beta, rm, rv = make(300, 96, 1)
    rm0, rv0 = rm.clone(), rv.clone()
    ref = F.relu(F.batch_norm(z, rm, rv, gamma, beta, False, 0.1, 1e-5))
    out, _ = dk.bn_act_fwd(z, gamma, beta, rm, rv, False, False, 0.2, SEED, SALT, torch.float32)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-5)
    assert torch.equal(rm, rm0) and torch.equal(rv, rv0)                 # eval never touches the running stats
    one, _ = dk.bn_act_fwd(z[:1], gamma, beta, rm, rv, True, True, 0.0, SEED, SALT, torch.float32)
    assert torch.equal(one, F.relu(z[:1]))                               # batch of one: BatchNorm skipped


def test_dropout_stream():
    m, width, p = 4000, 96, 0.2
    z, gamma, beta, rm, rv = make(m, width, 3)
    keep = dk.dropout_mask(SEED, SALT, (m, width), p, DEV)
    assert abs(float(keep.float().mean()) - (1 - p)) < 5e-3
    assert not torch.equal(keep, dk.dropout_mask(SEED + 1, SALT, (m, width), p, DEV))
    assert not torch.equal(keep, dk.dropout_mask(SEED, SALT + 1, (m, width), p, DEV))
    ref = F.relu(F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5)) * keep / (1 - p)
    out, _ = dk.bn_act_fwd(z, gamma, beta, rm, rv, True, False, p, SEED, SALT, torch.float32)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("m,width,p", [(37, 24, 0.0), (5000, 1024, 0.2), (300, 96, 0.5), (1, 8, 0.0)])
def test_bn_act_backward(m, width, p):
    z, gamma, beta, rm, rv = make(m, width, 5 + m)
    skip = m == 1
    d_out = torch.randn(m, width, device=DEV)
    keep = dk.dropout_mask(SEED, SALT, (m, width), p, DEV) if p > 0 else torch.ones(m, width, device=DEV)
    zr, gr, br = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = zr if skip else F.batch_norm(zr, None, None, gr, br, True, 0.1, 1e-5)
    (F.relu(y) * keep / (1 - p) * d_out).sum().backward()
    _, saved = dk.bn_act_fwd(z, gamma, beta, rm, rv, True, skip, p, SEED, SALT, None)
    dz, d_gamma, d_beta, d_bias = dk.bn_act_bwd(z, d_out, saved, skip, p, SEED, SALT, torch.float32)
    scale = float(zr.grad.abs().max())
    torch.testing.assert_close(dz, zr.grad, rtol=1e-4, atol=1e-5 * max(scale, 1))
    if skip:
        torch.testing.assert_close(d_bias, zr.grad.sum(0), rtol=1e-5, atol=1e-6)
    else:
        torch.testing.assert_close(d_gamma, gr.grad, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(d_beta, br.grad, rtol=1e-4, atol=1e-3)
        assert float(d_bias.abs().max()) == 0.0
    dz16, *_ = dk.bn_act_bwd(z, d_out, saved, skip, p, SEED, SALT, torch.bfloat16)
    assert torch.equal(dz16, dz.to(torch.bfloat16))


@pytest.mark.parametrize("m,width,ne,ng,p", [(500, 64, 4, 3, 0.0), (37, 8, 3, 2, 0.2), (1, 64, 4, 3, 0.0)])
def test_mmoe_mix(m, width, ne, ng, p):
    g = torch.Generator(device=DEV).manual_seed(m)
    z = torch.randn(m, ne * width, device=DEV, generator=g)
    saved = torch.stack([torch.zeros(ne * width, device=DEV), torch.ones(ne * width, device=DEV),
                         1 + 0.1 * torch.randn(ne * width, device=DEV, generator=g),
                         0.1 * torch.randn(ne * width, device=DEV, generator=g)])
    gate = torch.softmax(torch.randn(m, ng, ne, device=DEV, generator=g), dim=2)
    keep = dk.dropout_mask(SEED, SALT, (m, ne * width), p, DEV) if p > 0 else torch.ones(m, ne * width, device=DEV)
    zr, gr = z.clone().requires_grad_(True), gate.clone().requires_grad_(True)
    h_act = F.relu(zr * saved[2] + saved[3])
    h_act.retain_grad()
    h = (h_act * keep / (1 - p)).view(m, ne, width)
    ref = torch.einsum("bge,bew->bgw", gr, h)
    out = dk.mmoe_mix_fwd(z, saved, gate, ne, ng, p, SEED, SALT)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    d_out = torch.randn(m, ng, width, device=DEV, generator=g)
    (ref * d_out).sum().backward()
    d_h, d_gate = dk.mmoe_mix_bwd(z, saved, gate, d_out, ne, ng, p, SEED, SALT)
    torch.testing.assert_close(d_gate, gr.grad, rtol=1e-4, atol=1e-4)
    # d_h is the gradient w.r.t. the activated + dropped h (the BN backward applies relu'/dropout itself)
    ref_dh = torch.einsum("bge,bgw->bew", gate, d_out).reshape(m, ne * width)
    torch.testing.assert_close(d_h, ref_dh, rtol=1e-5, atol=1e-5)
