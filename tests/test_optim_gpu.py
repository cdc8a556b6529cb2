"""FusedAdam (C ABI: aread_adam_step) against torch.optim.Adam with the trainer's settings
(run.py:830-831).  Same arithmetic order, so the tolerance is fp32 contraction noise."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
optim = importlib.import_module("aread-multi-domain-recommendation_b200.optim")
DEV = "cuda:0"


def test_fused_adam_matches_torch_adam():
    gen = torch.Generator(device=DEV).manual_seed(0)
    shapes = [(50000, 32), (1, 288), (256, 288), (256,), (7,), (4097,), (64, 64), (3, 32)]
    ref = [torch.randn(*s, device=DEV, generator=gen).requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    kw = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    o_ref, o_mine = torch.optim.Adam(ref, **kw), optim.FusedAdam(mine, **kw)
    for step in range(6):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if (step + i) % 4 == 3:                   # parameters without gradient are skipped entirely
                a.grad = b.grad = None
                continue
            g = torch.randn(a.shape, device=DEV, generator=gen) * (10.0 ** ((i % 5) - 3))
            a.grad, b.grad = g.clone(), g.clone()
        o_ref.step()
        o_mine.step()
        for i, (a, b) in enumerate(zip(ref, mine)):
            torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-7, msg=f"param {i} step {step}")
    for a, b in zip(ref, mine):
        sa, sb = o_ref.state[a], o_mine.state[b]
        assert float(sa["step"]) == float(sb["step"])
        # torch forms m with lerp, the kernel with two fmas: where beta1*m and (1-beta1)*g cancel, the difference
        # is an ulp of the operands, not of the result
        torch.testing.assert_close(sb["exp_avg"], sa["exp_avg"], rtol=5e-6, atol=1e-6 * float(sa["exp_avg"].abs().max()))
        torch.testing.assert_close(sb["exp_avg_sq"], sa["exp_avg_sq"], rtol=5e-6, atol=1e-12)
    # state_dict round trip into a torch Adam keeps training identically
    sd = o_mine.state_dict()
    o_back = torch.optim.Adam(mine, **kw)
    o_back.load_state_dict(sd)
    assert len(o_back.state) == len(o_mine.state)


def test_folded_l2_matches_regulariser_in_the_loss():
    """FusedAdam.fold_l2: same weights as adding l2 * sum(w^2) to the loss, including parameters whose only
    gradient is the regulariser's (grad None otherwise)."""
    gen = torch.Generator(device=DEV).manual_seed(3)
    shapes = [(20000, 32), (256, 288), (256,), (64, 64), (5,)]
    l2 = [1e-5, 1e-5, 0.0, 2e-5, 0.0]
    a = [torch.randn(*s, device=DEV, generator=gen).requires_grad_(True) for s in shapes]
    b = [p.detach().clone().requires_grad_(True) for p in a]
    kw = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    o_a, o_b = torch.optim.Adam(a, **kw), optim.FusedAdam(b, **kw)
    o_b.fold_l2({p: c for p, c in zip(b, l2) if c > 0})
    for step in range(5):
        data = [None if (step + i) % 3 == 2 else torch.randn(s, device=DEV, generator=gen) * 1e-3
                for i, s in enumerate(shapes)]
        for p, q, g, c in zip(a, b, data, l2):
            reg = 2 * c * p.detach() if c > 0 else None
            p.grad = g.clone() + reg if (g is not None and reg is not None) else (g.clone() if g is not None else reg)
            q.grad = None if g is None else g.clone()
        o_a.step()
        o_b.step()
        for i, (p, q) in enumerate(zip(a, b)):
            torch.testing.assert_close(q, p, rtol=1e-6, atol=1e-7, msg=f"param {i} step {step}")
    with pytest.raises(ValueError):
        o_b.fold_l2({torch.nn.Parameter(torch.zeros(3, device=DEV)): 1e-5})
