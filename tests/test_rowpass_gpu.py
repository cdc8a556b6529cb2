"""Row pass kernels (C ABI: aread_rowpass_fwd / _bwd) against a plain torch fp32 reference that
evaluates the linear term, gate softmax, cross network and head dot products the way the reference
modules do (explicit cross-network states).  fp32 in both; tolerance = summation-order round-off."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
rp = importlib.import_module("aread-multi-domain-recommendation_b200.rowpass_ops")
DEV = "cuda:0"


def reference(x, w_lin, b_lin, w_gate, b_gate, w_cn, b_cn, w_out):
    lin = x @ w_lin.t() + b_lin                                    # [B, 1]
    gate = torch.softmax(torch.einsum("be,gke->bgk", x, w_gate) + b_gate, dim=2)
    c = x
    for k in range(w_cn.shape[0]):
        c = x * (c @ w_cn[k:k + 1].t()) + b_cn[k] + c
    head = c @ w_out.t()
    return lin.squeeze(1), gate, head


@pytest.mark.parametrize("m,e,ng,ne,nc,nh", [(300, 288, 3, 4, 3, 12), (1000, 736, 3, 4, 3, 12), (37, 40, 2, 3, 2, 8),
                                             (1, 288, 1, 4, 3, 2), (5000, 160, 3, 4, 0, 5), (65, 736, 0, 4, 3, 1)])
def test_rowpass_forward_backward(m, e, ng, ne, nc, nh):
    g = torch.Generator(device=DEV).manual_seed(m + e)
    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=DEV, generator=g) * scale).requires_grad_(True)
    x = rnd(m, e)
    w_lin, b_lin = rnd(1, e, scale=e ** -0.5), rnd(1, scale=0.1)
    w_gate, b_gate = rnd(ng, ne, e, scale=e ** -0.5), rnd(ng, ne, scale=0.1)
    w_cn, b_cn = rnd(nc, e, scale=e ** -0.5), rnd(nc, e, scale=0.05)
    w_out = rnd(nh, e, scale=e ** -0.5)
    ref = reference(x, w_lin, b_lin, w_gate, b_gate, w_cn, b_cn, w_out)

    w = torch.cat([w_lin, w_gate.reshape(ng * ne, e), w_cn, w_out], dim=0)
    beta = torch.zeros(e, device=DEV)
    kappas = []
    for k in range(nc):
        kappas.append((w_cn[k] * beta).sum().reshape(1))
        beta = beta + b_cn[k]
    offset = torch.cat([b_lin, b_gate.reshape(-1)] + kappas + [w_out @ beta])
    lin, gate, head, _ = rp.RowPass.apply(x, w, offset, (ng, ne, nc, nh))
    torch.testing.assert_close(lin, ref[0], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(gate, ref[1], rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(head, ref[2], rtol=5e-5, atol=5e-5)

    d = [torch.randn(t.shape, device=DEV, generator=g) for t in ref]
    params = [x, w_lin, b_lin, w_gate, b_gate, w_cn, b_cn, w_out]
    want = torch.autograd.grad(sum((r * dd).sum() for r, dd in zip(ref, d)), params, allow_unused=True)
    got = torch.autograd.grad((lin * d[0]).sum() + (gate * d[1]).sum() + (head * d[2]).sum(), params, allow_unused=True)
    for name, a, b in zip("x w_lin b_lin w_gate b_gate w_cn b_cn w_out".split(), got, want):
        if b is None or b.numel() == 0:
            continue
        a = torch.zeros_like(b) if a is None else a
        scale = float(b.abs().max()) + 1e-6
        assert float((a - b).abs().max()) <= 2e-4 * scale + 1e-6, (name, float((a - b).abs().max()), scale)


def test_tensor_core_row_pass_matches_cuda_core_row_pass(monkeypatch):
    """The row pass on the tensor cores (split operands) with the HEI gate logits riding along, against the all-fp32
    CUDA-core row pass + separate gate Linears, on the same bf16-expert model: probabilities to 1e-4, every gradient
    family to 2e-2 (the forward products differ by ~2^-17 relative; the gradients behind the towers amplify that)."""
    import importlib
    import numpy as np
    from oracle import aread_torch as O
    from oracle import synth
    from tests._models import build_model
    from tests._util import family_errors, load_golden
    fused = importlib.import_module("aread-multi-domain-recommendation_b200.fused")
    fx = load_golden("ali_small")
    spec = O.Spec(**fx["spec"])
    x, y = synth.random_batch(spec, 2048, seed=5, domain=fx["domain"])
    results = []
    for tc in (True, False):
        monkeypatch.setattr(fused, "TC_ROWPASS", tc)
        monkeypatch.setattr(fused, "USE_GRAPHS", False)
        model = build_model(spec, "cuda:0", dropout=0.0).train()
        model.expert_precision = "bf16"
        np.random.seed(3)
        mask = model.generate_mask("rand", 0, init_active_percent=0.5)
        preds = model(x.to("cuda:0"), mode="domain_mask_bagging", current_mask=mask, tmp_memory_gate_value=True)
        loss = model.bagging_loss(preds, y.to("cuda:0"))
        model.zero_grad()
        loss.backward()
        means = [torch.stack(model.tmp_tower_gate_values[l], dim=1).cpu() for l in range(1, spec.n_level)]
        results.append((preds.detach().cpu(), means,
                        {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}))
    (p_tc, m_tc, g_tc), (p_cc, m_cc, g_cc) = results
    assert g_tc.keys() == g_cc.keys()
    errs = family_errors([(k, g_tc[k].reshape(-1), g_cc[k].reshape(-1)) for k in g_cc])
    report = {"prob": float((p_tc - p_cc).abs().max()),
              "gate_means": max(float((a - b).abs().max()) for a, b in zip(m_tc, m_cc)),
              "grads": {f: round(e, 5) for f, (e, _) in errs.items()}}
    assert report["prob"] <= 1e-4, report
    assert report["gate_means"] <= 2e-6, report            # what HEMP thresholds
    assert all(e <= 2e-2 for e in report["grads"].values()), report
