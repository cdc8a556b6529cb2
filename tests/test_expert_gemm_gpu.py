"""aread_expert_gemm (tcgen05 GEMM with the BatchNorm bookkeeping in the epilogue) and its companions against a
plain torch fp32 reference of the same ops on the same bf16-rounded operands.

Tolerances: accumulators are fp32 in TMEM, so GEMM values differ from torch by accumulation order only; bf16 outputs
add one rounding (rel 2^-8); column sums over m rows are compared with rel 1e-4 of the column's absolute sum."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")
DEV = "cuda:0"

# m, n, k, groups, a_group_cols  (n % 64 == 0 for the fused epilogues)
FWD_SHAPES = [(300, 256, 288, 4, 0), (257, 256, 736, 4, 0), (1000, 128, 256, 4, 256), (129, 64, 128, 4, 128),
              (4096, 256, 736, 4, 0), (1, 64, 128, 2, 128), (65, 128, 64, 3, 64)]


def _operands(m, n, k, groups, agc, seed=0):
    gen = torch.Generator(device=DEV).manual_seed(seed + m * 7 + n)
    a_cols = k if agc == 0 else agc * groups
    a = torch.randn(m, a_cols, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(groups * n, k, device=DEV, generator=gen) / k ** 0.5).to(torch.bfloat16)
    return a, w, gen


def _ref_fwd(a, w, n, k, groups, agc):
    out = torch.empty(a.shape[0], groups * n, dtype=torch.float32, device=DEV)
    for g in range(groups):
        out[:, g * n:(g + 1) * n] = a[:, g * agc:g * agc + k].float() @ w[g * n:(g + 1) * n, :k].float().t()
    return out


def _bf16_close(got, ref, what):
    got, ref = got.float(), ref.float()
    tol = ref.abs() * 2.0 ** -8 + 1e-5
    assert bool(((got - ref).abs() <= tol).all()), f"{what}: {float(((got - ref).abs() - tol).max()):.3e} over tolerance"


@pytest.mark.parametrize("m,n,k,groups,agc", FWD_SHAPES)
def test_stats_epilogue(m, n, k, groups, agc):
    a, w, _ = _operands(m, n, k, groups, agc)
    z, partial = dk.expert_linear_stats(a, w, n, k, groups, agc)
    torch.cuda.synchronize()
    ref = _ref_fwd(a, w, n, k, groups, agc)
    _bf16_close(z, ref, "z16")
    s = partial[:, 0, :].sum(dim=0)
    q = partial[:, 1, :].sum(dim=0)
    assert partial.shape[0] == (m + 127) // 128
    torch.testing.assert_close(s, ref.sum(dim=0), rtol=0, atol=1e-4 * float(ref.abs().sum(dim=0).max()) + 1e-5)
    torch.testing.assert_close(q, (ref * ref).sum(dim=0), rtol=1e-4, atol=1e-5)
    # per-tile partials: rows of tile t only
    t_last = partial.shape[0] - 1
    torch.testing.assert_close(partial[t_last, 0], ref[t_last * 128:].sum(dim=0), rtol=0,
                               atol=1e-4 * float(ref[t_last * 128:].abs().sum(dim=0).max()) + 1e-5)
    # deterministic
    z2, partial2 = dk.expert_linear_stats(a, w, n, k, groups, agc)
    assert torch.equal(z2, z) and torch.equal(partial2, partial)


@pytest.mark.parametrize("m,n,k,groups,agc", FWD_SHAPES[:5])
@pytest.mark.parametrize("training", [True, False])
def test_finalize_and_activation(m, n, k, groups, agc, training):
    a, w, gen = _operands(m, n, k, groups, agc, seed=3)
    width = groups * n
    bias = torch.randn(width, device=DEV, generator=gen)
    gamma = 1 + 0.1 * torch.randn(width, device=DEV, generator=gen)
    beta = 0.1 * torch.randn(width, device=DEV, generator=gen)
    rm0 = 0.1 * torch.randn(width, device=DEV, generator=gen)
    rv0 = 0.5 + torch.rand(width, device=DEV, generator=gen)
    rm, rv = rm0.clone(), rv0.clone()
    z, partial = dk.expert_linear_stats(a, w, n, k, groups, agc)
    saved = dk.expert_bn_finalize(partial, m, width, bias, gamma, beta, rm, rv, training, False)
    h = dk.bn16_fwd(z, saved, training, 0.0, 0, 0)
    torch.cuda.synchronize()
    zf = z.float()                      # what the kernels normalise: the stored bf16 pre-activation (bias-free)
    full = _ref_fwd(a, w, n, k, groups, agc) + bias
    bn = torch.nn.functional.batch_norm(full, rm0.clone(), rv0.clone(), gamma, beta, training, 0.1, 1e-5)
    ref_h = torch.relu(bn)
    # normalisation of the bf16-stored accumulator vs BatchNorm of the fp32 Linear output: rounding of z only
    assert float((h.float() - ref_h).abs().max()) <= 2e-2 * float(ref_h.abs().max()) + 1e-3
    assert float((h.float() - ref_h).abs().mean()) <= 3e-3
    if training and m > 1:
        ref_rm, ref_rv = rm0.clone(), rv0.clone()
        torch.nn.functional.batch_norm(full, ref_rm, ref_rv, gamma, beta, True, 0.1, 1e-5)
        torch.testing.assert_close(rm, ref_rm, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(rv, ref_rv, rtol=1e-3, atol=1e-5)
        mean_a = full.mean(dim=0) - bias
        torch.testing.assert_close(saved[0], mean_a, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(saved[1], 1 / torch.sqrt(full.var(dim=0, unbiased=False) + 1e-5), rtol=1e-3, atol=1e-5)
    else:
        assert torch.equal(rm, rm0) and torch.equal(rv, rv0)
        # inference: the same numbers from the epilogue that folds BatchNorm + ReLU
        folded = dk.expert_bn_finalize(None, m, width, bias, gamma, beta, rm, rv, False, False)
        h2 = dk.expert_linear_act(a, w, n, k, groups, agc, folded)
        _bf16_close(h2, ref_h, "ACT epilogue")
    # exact elementwise check of bn16_fwd on its own inputs
    want = torch.relu(zf * saved[2] + saved[3])              # the kernel uses one fma: up to a bf16 ulp apart
    assert float((h.float() - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max()) + 1e-6


def test_bn_skip_is_identity_plus_bias():
    m, n, k, groups, agc = 1, 64, 128, 2, 128
    a, w, gen = _operands(m, n, k, groups, agc)
    bias = torch.randn(groups * n, device=DEV, generator=gen)
    z, partial = dk.expert_linear_stats(a, w, n, k, groups, agc)
    saved = dk.expert_bn_finalize(partial, m, groups * n, bias, None, None, None, None, True, True)
    h = dk.bn16_fwd(z, saved, True, 0.0, 0, 0)
    assert torch.equal(h, torch.relu(z.float() + bias).to(torch.bfloat16))


@pytest.mark.parametrize("m,n_out,k,groups", [(300, 256, 128, 4), (129, 128, 64, 4), (1000, 64, 64, 3), (4096, 256, 128, 4)])
@pytest.mark.parametrize("p", [0.0, 0.2])
def test_bn_bwd_epilogue_and_apply(m, n_out, k, groups, p):
    """dA = dZ . W with the weight read in place as a k-by-n operand; the epilogue masks the gradient with the ReLU /
    dropout pattern of the layer below and accumulates sum(dy), sum(dy * xhat)."""
    gen = torch.Generator(device=DEV).manual_seed(m + n_out)
    width = groups * n_out
    dz = torch.randn(m, groups * k, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(groups * k, n_out, device=DEV, generator=gen) / k ** 0.5).to(torch.bfloat16)   # [G*N, K_in]
    z_prev = torch.randn(m, width, device=DEV, generator=gen).to(torch.bfloat16)
    saved = torch.empty(4, width, device=DEV)
    saved[0] = 0.1 * torch.randn(width, device=DEV, generator=gen)            # mean
    saved[1] = 0.5 + torch.rand(width, device=DEV, generator=gen)             # rstd
    saved[2] = saved[1] * (1 + 0.1 * torch.randn(width, device=DEV, generator=gen))   # scale
    saved[3] = 0.1 * torch.randn(width, device=DEV, generator=gen)            # shift
    seed, salt = 1234567, 0x1001
    dy, partial = dk.expert_dgrad_bn_bwd(dz, w, n_out, k, groups, z_prev, saved, p, salt, seed)
    torch.cuda.synchronize()
    acc = torch.empty(m, width, device=DEV)
    for g in range(groups):
        acc[:, g * n_out:(g + 1) * n_out] = dz[:, g * k:(g + 1) * k].float() @ w[g * k:(g + 1) * k].float()
    zf = z_prev.float()
    y = zf * saved[2] + saved[3]
    keep = dk.dropout_mask(seed, salt, (m, width), p, DEV) if p > 0 else torch.ones(m, width, dtype=torch.bool, device=DEV)
    ref_dy = torch.where((y > 0) & keep, acc / (1 - p), torch.zeros_like(acc))
    _bf16_close(dy, ref_dy, "dy16")
    xhat = (zf - saved[0]) * saved[1]
    s1, s2 = partial[:, 0].sum(dim=0), partial[:, 1].sum(dim=0)
    scale1 = float(ref_dy.abs().sum(dim=0).max())
    torch.testing.assert_close(s1, ref_dy.sum(dim=0), rtol=0, atol=1e-4 * scale1 + 1e-5)
    torch.testing.assert_close(s2, (ref_dy * xhat).sum(dim=0), rtol=0, atol=1e-4 * float((ref_dy * xhat).abs().sum(dim=0).max()) + 1e-5)
    coef, grads = dk.expert_bn_bwd_finalize(partial, m, width, False)
    torch.testing.assert_close(grads[0], s2, rtol=1e-4, atol=1e-4 * scale1)       # same partials, another order
    torch.testing.assert_close(grads[1], s1, rtol=1e-4, atol=1e-4 * scale1)
    assert not grads[2].any()
    dzl = dk.bn16_bwd(z_prev, dy, saved, coef, False)
    want = (saved[2] * (dy.float() - coef[0] - xhat * coef[1])).to(torch.bfloat16)
    assert float((dzl.float() - want.float()).abs().max()) <= 2.0 ** -7 * float(want.float().abs().max()) + 1e-6


@pytest.mark.parametrize("m,n_out,k", [(300, 736, 1024), (129, 288, 1024), (2048, 736, 1024), (5, 64, 64)])
def test_plain_dgrad_k_by_n(m, n_out, k):
    gen = torch.Generator(device=DEV).manual_seed(m)
    dz = torch.randn(m, k, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(k, n_out, device=DEV, generator=gen) / k ** 0.5).to(torch.bfloat16)
    got = dk.expert_dgrad_plain(dz, w, n_out, k)
    torch.cuda.synchronize()
    ref = dz.float() @ w.float()
    tol = 2e-5 * (dz.float().abs() @ w.float().abs()) + 1e-6
    assert bool(((got - ref).abs() <= tol).all()), float(((got - ref).abs() - tol).max())


def test_weight_cast_one_launch():
    flats = [torch.randn(4, 256, 736, device=DEV), torch.randn(4, 128, 256, device=DEV), torch.randn(4, 64, 128, device=DEV)]
    wc = dk.WeightCast(flats)
    outs = wc.run()
    for f, o in zip(flats, outs):
        assert torch.equal(o, f.to(torch.bfloat16))
    flats[1].mul_(3)
    assert torch.equal(wc.run()[1], flats[1].to(torch.bfloat16))


@pytest.mark.parametrize("m,width", [(300, 1024), (4096, 512), (129, 64), (1000, 256)])
@pytest.mark.parametrize("p", [0.0, 0.2])
def test_unfused_bn_backward_passes(m, width, p):
    """bn16_fwd's pass bits, bn16_bwd_stats and bn16_bwd(raw=True) -- the default backward of the bf16 experts -- against
    torch on the same inputs; with and without the bit map (the mask is then rebuilt from the dropout stream)."""
    gen = torch.Generator(device=DEV).manual_seed(m + width)
    z = torch.randn(m, width, device=DEV, generator=gen).to(torch.bfloat16)
    d_h = torch.randn(m, width, device=DEV, generator=gen).to(torch.bfloat16)
    saved = torch.empty(4, width, device=DEV)
    saved[0] = 0.1 * torch.randn(width, device=DEV, generator=gen)
    saved[1] = 0.5 + torch.rand(width, device=DEV, generator=gen)
    saved[2] = saved[1] * (1 + 0.1 * torch.randn(width, device=DEV, generator=gen))
    saved[3] = 0.1 * torch.randn(width, device=DEV, generator=gen)
    seed, salt = 987654321, 0x1002
    h, bits = dk.bn16_fwd(z, saved, True, p, seed, salt, want_bits=True)
    zf = z.float()
    y = zf * saved[2] + saved[3]
    keep = dk.dropout_mask(seed, salt, (m, width), p, DEV) if p > 0 else torch.ones(m, width, dtype=torch.bool, device=DEV)
    passed = (y > 0) & keep
    want_h = torch.where(passed, y / (1 - p), torch.zeros_like(y))
    assert float((h.float() - want_h).abs().max()) <= 2.0 ** -7 * float(want_h.abs().max()) + 1e-6
    unpacked = ((bits.view(m, width // 8, 1).int() >> torch.arange(8, device=DEV).view(1, 1, 8)) & 1).bool().view(m, width)
    assert torch.equal(unpacked, h != 0) or torch.equal(unpacked, passed)
    if p > 0:
        assert abs(float(keep.float().mean()) - (1 - p)) < 5e-3
    dy = torch.where(unpacked, d_h.float() / (1 - p), torch.zeros_like(y))
    xhat = (zf - saved[0]) * saved[1]
    for use_bits in (True, False):
        partial = dk.bn16_bwd_stats(z, d_h, saved, False, p, salt, seed, bits=bits if use_bits else None)
        s1, s2 = partial[:, 0].sum(dim=0), partial[:, 1].sum(dim=0)
        torch.testing.assert_close(s1, dy.sum(dim=0), rtol=0, atol=1e-4 * float(dy.abs().sum(dim=0).max()) + 1e-5)
        torch.testing.assert_close(s2, (dy * xhat).sum(dim=0), rtol=0, atol=1e-4 * float((dy * xhat).abs().sum(dim=0).max()) + 1e-5)
        coef, grads = dk.expert_bn_bwd_finalize(partial, m, width, False)
        dz = dk.bn16_bwd(z, d_h, saved, coef, False, raw=True, p=p, salt=salt, seed=seed, bits=bits if use_bits else None)
        want = saved[2] * (dy - coef[0] - xhat * coef[1])
        assert float((dz.float() - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max()) + 1e-5
