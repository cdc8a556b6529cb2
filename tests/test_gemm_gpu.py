"""tcgen05 grouped Linear (C ABI: aread_grouped_linear_bf16) against a plain torch fp32 reference of
the same op on the same bf16-rounded operands.  Tolerance: fp32 accumulation-order noise only
(|d| <= 2e-5 * sum_k |a||w| + 1e-6); the bf16 output adds one rounding (rel 2^-8)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")
DEV = "cuda:0"

SHAPES = [
    # m,    n,    k,   groups, a_group_cols
    (300, 1024, 288, 1, 0),       # expert layer 1, Amazon-shaped E (K tail: 288 = 4.5 x 64)
    (257, 1024, 736, 1, 0),       # expert layer 1, AliCCP-shaped E
    (1000, 128, 256, 4, 256),     # expert layer 2: block diagonal
    (129, 64, 128, 4, 128),       # expert layer 3
    (37, 32, 40, 3, 0),           # tiny: n < tile, k < tile, shared A
    (5, 16, 32, 3, 32),           # tiny grouped
    (1, 8, 16, 2, 16),            # single row
    (4096, 256, 1024, 1, 0),      # many k blocks (pipeline wrap-around), many m tiles
]


def reference(a, w, bias, n, k, groups, a_group_cols, mask):
    m = a.shape[0]
    out = torch.zeros(m, groups * n, dtype=torch.float32, device=a.device)
    for g in range(groups):
        if not (mask >> g) & 1:
            continue
        ag = a[:, g * a_group_cols:g * a_group_cols + k].float()
        wg = w[g * n:(g + 1) * n, :k].float()
        out[:, g * n:(g + 1) * n] = ag @ wg.t() + (bias[g * n:(g + 1) * n] if bias is not None else 0)
    return out


def bound(a, w, n, k, groups, a_group_cols):
    m = a.shape[0]
    out = torch.zeros(m, groups * n, dtype=torch.float32, device=a.device)
    for g in range(groups):
        ag = a[:, g * a_group_cols:g * a_group_cols + k].float().abs()
        out[:, g * n:(g + 1) * n] = ag @ w[g * n:(g + 1) * n, :k].float().abs().t()
    return out


@pytest.mark.parametrize("m,n,k,groups,agc", SHAPES)
@pytest.mark.parametrize("use_bias", [False, True])
def test_grouped_linear_fp32_out(m, n, k, groups, agc, use_bias):
    gen = torch.Generator(device=DEV).manual_seed(m * 7 + n)
    a_cols = k if agc == 0 else agc * groups
    a = torch.randn(m, a_cols, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(groups * n, k, device=DEV, generator=gen) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn(groups * n, device=DEV, generator=gen) if use_bias else None
    full = (1 << groups) - 1
    got = dk.grouped_linear(a, w, bias, n, k, groups, agc)
    torch.cuda.synchronize()
    ref = reference(a, w, bias, n, k, groups, agc, full)
    tol = 2e-5 * bound(a, w, n, k, groups, agc) + 1e-6
    assert bool(((got - ref).abs() <= tol).all()), float(((got - ref).abs() - tol).max())


def test_group_mask_skips_groups():
    m, n, k, groups = 500, 128, 256, 4
    gen = torch.Generator(device=DEV).manual_seed(3)
    a = torch.randn(m, k * groups, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(groups * n, k, device=DEV, generator=gen) / 16).to(torch.bfloat16)
    for mask in (0b0101, 0b1000, 0b0000):
        out = torch.full((m, groups * n), 7.0, device=DEV)
        dk.grouped_linear(a, w, None, n, k, groups, k, group_mask=mask, out=out)
        ref = reference(a, w, None, n, k, groups, k, mask)
        for g in range(groups):
            cols = slice(g * n, (g + 1) * n)
            if (mask >> g) & 1:
                torch.testing.assert_close(out[:, cols], ref[:, cols], rtol=1e-4, atol=1e-4)
            else:
                assert bool((out[:, cols] == 7.0).all())          # untouched


def test_bf16_output_and_strided_operands():
    m, n, k = 777, 128, 256
    gen = torch.Generator(device=DEV).manual_seed(5)
    big = torch.randn(m, 1024 + 8, device=DEV, generator=gen).to(torch.bfloat16)
    a = big[:, 8:8 + 1024]                                          # lda = 1032, 16-byte aligned start
    w = (torch.randn(4 * n, k, device=DEV, generator=gen) / 16).to(torch.bfloat16)
    got = dk.grouped_linear(a, w, None, n, k, 4, 256, out_dtype=torch.bfloat16)
    ref = reference(a, w, None, n, k, 4, 256, 0b1111)
    torch.testing.assert_close(got.float(), ref, rtol=2 ** -7, atol=1e-3)
    assert got.dtype == torch.bfloat16


def test_repeated_launches_are_bit_identical():
    m, n, k = 2048, 1024, 736
    gen = torch.Generator(device=DEV).manual_seed(9)
    a = torch.randn(m, k, device=DEV, generator=gen).to(torch.bfloat16)
    w = (torch.randn(n, k, device=DEV, generator=gen) / 27).to(torch.bfloat16)
    first = dk.grouped_linear(a, w, None, n, k, 1)
    for _ in range(3):
        assert torch.equal(first, dk.grouped_linear(a, w, None, n, k, 1))


WGRAD_SHAPES = [
    # m,     n,    k,  groups, a_group_cols
    (3000, 1024, 288, 1, 0),      # expert layer 1 (n > 128: several output-feature tiles; k tail)
    (1000, 256, 736, 4, 0),       # four experts sharing the input
    (5000, 128, 256, 4, 256),     # expert layer 2
    (777, 64, 128, 4, 128),       # expert layer 3 (n < 128)
    (37, 32, 40, 3, 0),           # tiny
    (1, 8, 16, 2, 16),            # one sample
    (70000, 128, 64, 2, 64),      # many sample blocks -> split over CTAs
]


@pytest.mark.parametrize("m,n,k,groups,agc", WGRAD_SHAPES)
def test_grouped_wgrad(m, n, k, groups, agc):
    gen = torch.Generator(device=DEV).manual_seed(m + n + k)
    a_cols = k if agc == 0 else agc * groups
    a = torch.randn(m, a_cols, device=DEV, generator=gen).to(torch.bfloat16)
    dz = (torch.randn(m, groups * n, device=DEV, generator=gen) / m ** 0.5).to(torch.bfloat16)
    got = dk.grouped_wgrad(dz, a, n, k, groups, agc)
    torch.cuda.synchronize()
    for g in range(groups):
        ag = a[:, g * agc:g * agc + k].double()
        dzg = dz[:, g * n:(g + 1) * n].double()
        ref = dzg.t() @ ag
        tol = 3e-5 * (dzg.abs().t() @ ag.abs()) + 1e-6
        err = (got[g * n:(g + 1) * n].double() - ref).abs()
        assert bool((err <= tol).all()), (g, float((err - tol).max()))


def test_grouped_wgrad_mask_and_determinism():
    m, n, k, groups = 9000, 128, 256, 4
    gen = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(m, k * groups, device=DEV, generator=gen).to(torch.bfloat16)
    dz = torch.randn(m, groups * n, device=DEV, generator=gen).to(torch.bfloat16)
    out = torch.full((groups * n, k), 3.0, device=DEV)
    dk.grouped_wgrad(dz, a, n, k, groups, k, group_mask=0b0110, out=out)
    assert bool((out[:n] == 3.0).all()) and bool((out[3 * n:] == 3.0).all())
    full = dk.grouped_wgrad(dz, a, n, k, groups, k)
    # the sample split depends on the number of active groups, so masked vs full differ by round-off only
    torch.testing.assert_close(out[n:3 * n], full[n:3 * n], rtol=1e-5, atol=1e-3)
    assert torch.equal(full, dk.grouped_wgrad(dz, a, n, k, groups, k))
    again = torch.full((groups * n, k), 3.0, device=DEV)
    dk.grouped_wgrad(dz, a, n, k, groups, k, group_mask=0b0110, out=again)
    assert torch.equal(out, again)


@pytest.mark.parametrize("m,n,k,groups,agc", [(300, 1024, 288, 1, 0), (1000, 128, 256, 4, 256), (37, 32, 40, 3, 0)])
def test_split_precision_linear_is_fp32_grade(m, n, k, groups, agc):
    """hi/lo split operands, three passes: error ~2^-16 of sum |a||w| instead of bf16's 2^-8."""
    gen = torch.Generator(device=DEV).manual_seed(m + k)
    a_cols = k if agc == 0 else agc * groups
    a32 = torch.randn(m, a_cols, device=DEV, generator=gen)
    w32 = torch.randn(groups * n, k, device=DEV, generator=gen) / k ** 0.5
    (a_hi, a_lo), (w_hi, w_lo) = dk.split_bf16(a32), dk.split_bf16(w32)
    got = dk.grouped_linear(a_hi, w_hi, None, n, k, groups, agc, a_lo=a_lo, w_lo=w_lo)
    one_pass = dk.grouped_linear(a_hi, w_hi, None, n, k, groups, agc)
    for g in range(groups):
        ag = a32[:, g * agc:g * agc + k].double()
        wg = w32[g * n:(g + 1) * n].double()
        ref = ag @ wg.t()
        bound = ag.abs() @ wg.abs().t()
        err3 = (got[:, g * n:(g + 1) * n].double() - ref).abs()
        err1 = (one_pass[:, g * n:(g + 1) * n].double() - ref).abs()
        assert bool((err3 <= 4e-5 * bound + 1e-6).all()), float((err3 / (bound + 1e-9)).max())
        assert float(err3.mean()) < 0.02 * float(err1.mean())          # two orders of magnitude better than one pass


def test_split_precision_wgrad_is_fp32_grade():
    m, n, k, groups = 5000, 128, 256, 4
    gen = torch.Generator(device=DEV).manual_seed(21)
    a32 = torch.randn(m, k * groups, device=DEV, generator=gen)
    dz32 = torch.randn(m, groups * n, device=DEV, generator=gen) / m ** 0.5
    (a_hi, a_lo), (z_hi, z_lo) = dk.split_bf16(a32), dk.split_bf16(dz32)
    got = dk.grouped_wgrad(z_hi, a_hi, n, k, groups, k, dz_lo=z_lo, a_lo=a_lo)
    for g in range(groups):
        ag = a32[:, g * k:(g + 1) * k].double()
        zg = dz32[:, g * n:(g + 1) * n].double()
        ref = zg.t() @ ag
        bound = zg.abs().t() @ ag.abs()
        err = (got[g * n:(g + 1) * n].double() - ref).abs()
        assert bool((err <= 4e-5 * bound + 1e-6).all()), float((err / (bound + 1e-9)).max())
