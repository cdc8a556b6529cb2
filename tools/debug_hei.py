"""Mismatch report of the HEI tower-layer kernels against autograd (development aid; same set-up as tests/test_hei_gpu.py)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ho = importlib.import_module("aread-multi-domain-recommendation_b200.hei_ops")
dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")
from tests.test_hei_gpu import _ref_layer  # noqa: E402

DEV = "cuda:0"


def report(name, got, want, rtol=1e-3, atol=2e-4):
    bad = (got - want).abs() > atol + rtol * want.abs()
    n = int(bad.sum())
    print(f"{name}: shape {tuple(got.shape)} mismatches {n} max|d| {float((got - want).abs().max()):.3e}")
    if n:
        idx = bad.nonzero()
        print("   first:", idx[:6].tolist(), " last:", idx[-3:].tolist())
        for d in range(idx.shape[1]):
            u = torch.unique(idx[:, d])
            print(f"   dim {d}: {len(u)} distinct, e.g. {u[:12].tolist()}")


def run(m, G, dims, p):
    print(f"==== m={m} G={G} dims={dims} p={p}")
    K, N1, N2 = dims
    gen = torch.Generator(device=DEV).manual_seed(m + G)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    x = rnd(m, G, K).requires_grad_(True)
    w1, b1 = (0.3 * rnd(G, N1, K)).requires_grad_(True), rnd(G, N1).requires_grad_(True)
    g1, be1 = (1 + 0.1 * rnd(G, N1)).requires_grad_(True), (0.1 * rnd(G, N1)).requires_grad_(True)
    w2, b2 = (0.3 * rnd(G, N2, N1)).requires_grad_(True), rnd(G, N2).requires_grad_(True)
    g2, be2 = (1 + 0.1 * rnd(G, N2)).requires_grad_(True), (0.1 * rnd(G, N2)).requires_grad_(True)
    d_u = rnd(m, G, N2)
    seed, salt1, salt2 = 1234567, 0x2001, 0x2002
    rm1, rv1 = torch.zeros(G * N1, device=DEV), torch.ones(G * N1, device=DEV)
    rm2, rv2 = torch.zeros(G * N2, device=DEV), torch.ones(G * N2, device=DEV)
    mask1 = dk.dropout_mask(seed, salt1, (m, G, N1), p, DEV).float()
    mask2 = dk.dropout_mask(seed, salt2, (m, G, N2), p, DEV).float()
    with torch.no_grad():
        src = x.detach().reshape(m, G * K)
        z1, s1 = ho.layer_fwd(src, None, 0, w1.detach(), b1.detach(), g1.detach().reshape(-1), be1.detach().reshape(-1),
                              rm1, rv1, G, K, N1, True, False, p, seed)
        z2, s2 = ho.layer_fwd(z1, s1, salt1, w2.detach(), b2.detach(), g2.detach().reshape(-1),
                              be2.detach().reshape(-1), rm2, rv2, G, N1, N2, True, False, p, seed)
        y1_k = torch.addcmul(s1[3].expand_as(z1), z1, s1[2].expand_as(z1)).view(m, G, N1)
        y2_k = torch.addcmul(s2[3].expand_as(z2), z2, s2[2].expand_as(z2)).view(m, G, N2)
    z1_ref, a1_ref = _ref_layer(x, w1, b1, g1, be1, mask1, p, False, y1_k)
    a1_ref.retain_grad()
    z2_ref, u_ref = _ref_layer(a1_ref, w2, b2, g2, be2, mask2, p, False, y2_k)
    (u_ref * d_u).sum().backward()
    with torch.no_grad():
        report("z1", z1.view(m, G, N1), z1_ref)
        report("z2", z2.view(m, G, N2), z2_ref)
        d_out = d_u.reshape(m, G * N2).contiguous()
        coef, g3 = ho.bn_bwd_coef(z2, d_out, s2, False, p, seed, salt2)
        d_a1, d_w2, coef1, g3_1 = ho.layer_bwd(z2, d_out, s2, coef, p, salt2, seed, False, z1, s1, salt1, w2.detach(), G, N1, N2)
        d_x, d_w1, _, _ = ho.layer_bwd(z1, d_a1, s1, coef1, p, salt1, seed, False, src, None, 0, w1.detach(), G, K, N1)
        torch.cuda.synchronize()
        report("d_a1", d_a1.view(m, G, N1), a1_ref.grad)
        report("d_w2", d_w2, w2.grad, atol=2e-4 * m ** 0.5)
        report("coef1", coef1, coef1, atol=1.0)
        report("d_gamma1", g3_1[0].view(G, N1), g1.grad, atol=2e-4 * m ** 0.5)
        report("d_beta1", g3_1[1].view(G, N1), be1.grad, atol=2e-4 * m ** 0.5)
        report("d_w1", d_w1, w1.grad, atol=2e-4 * m ** 0.5)
        report("d_x", d_x.view(m, G, K), x.grad)


if __name__ == "__main__":
    for cfg in [(1000, 3, (64, 64, 32), 0.0), (70000, 2, (16, 16, 8), 0.0), (70000, 4, (16, 16, 8), 0.0),
                (40000, 3, (64, 64, 32), 0.2), (33000, 2, (16, 16, 8), 0.0), (17000, 2, (16, 16, 8), 0.0)]:
        run(*cfg)
