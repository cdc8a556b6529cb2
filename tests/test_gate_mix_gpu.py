"""Gate mixing of one HEI level (C ABI: aread_gate_mix; model/aread.py:300-306, 313-318): softmax of the gate logits,
renormalisation over the edges the HEMP mask keeps, weighted sum of the previous level's ACTIVE tower outputs -- forward
and backward against the same arithmetic written with torch ops and differentiated by autograd.  Covers the tiled
kernels (rows of u read once per tile), logits that live as a column slice of a wider matrix with a per-logit
constant (the row-pass layout), pruned previous towers, ragged last tiles and tower counts that do not divide the CTA."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
fused = importlib.import_module("aread-multi-domain-recommendation_b200.fused")


def _reference(logits, offset, edges, prev_slot, u_prev):
    lg = logits if offset is None else logits + offset
    s = torch.softmax(lg, dim=-1)                                             # [B, NT, NP]
    if edges is not None:
        r = s * edges
        r = r / (r.sum(dim=-1, keepdim=True) + 1e-8)
        sm = s * edges
    else:
        r, sm = s, s
    B, NT, NP = lg.shape
    W = u_prev.shape[2]
    full = torch.zeros(B, NP, W, device=lg.device)
    cols = [j for j in range(NP) if prev_slot[j] >= 0]
    full = full.index_copy(1, torch.tensor(cols, device=lg.device), u_prev[:, [prev_slot[j] for j in cols]])
    return torch.einsum("btj,bjw->btw", r, full), sm


@pytest.mark.parametrize("m,NT,NP,active,W,masked,strided", [
    (1000, 4, 3, [0, 1, 2], 32, False, False), (4099, 8, 6, [0, 2, 3, 5], 16, True, True),
    (777, 5, 6, [1, 4], 16, True, False), (65, 12, 6, [0, 1, 2, 3, 4, 5], 16, True, True),
    (3, 1, 3, [2], 32, False, False), (2049, 3, 3, [0, 2], 64, True, True), (513, 7, 8, [0, 3, 7], 8, True, False)])
def test_gate_mix_matches_autograd(m, NT, NP, active, W, masked, strided):
    gen = torch.Generator(device=DEV).manual_seed(m + NT)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    NA = len(active)
    prev_slot = [-1] * NP
    for slot, j in enumerate(active):
        prev_slot[j] = slot
    edges = None
    if masked:       # edges only from active previous towers, at least one per tower
        edges = torch.zeros(NT, NP, device=DEV)
        for t in range(NT):
            keep = [j for k, j in enumerate(active) if (t + k) % 2 == 0] or [active[0]]
            edges[t, keep] = 1.0
    u_prev = rnd(m, NA, W).requires_grad_(True)
    offset = 0.3 * rnd(NT * NP) if strided else None
    if strided:      # logits as columns 5 .. 5 + NT * NP of a wider matrix
        wide = rnd(m, NT * NP + 11)
        logits2d = wide[:, 5:5 + NT * NP]
        lg_ref = logits2d.reshape(m, NT, NP).clone().requires_grad_(True)
        off_ref = offset.view(NT, NP)
        logits_arg = logits2d
    else:
        logits_arg = rnd(m, NT, NP)
        lg_ref = logits_arg.clone().requires_grad_(True)
        off_ref = None
    d_out = rnd(m, NT, W)
    out_ref, sm_ref = _reference(lg_ref, off_ref, edges, prev_slot, u_prev)
    (out_ref * d_out).sum().backward()

    slot_t = torch.tensor(prev_slot, dtype=torch.int32, device=DEV)
    slot_tower = torch.tensor(active, dtype=torch.int32, device=DEV)
    with torch.no_grad():
        out, sm = fused.gate_mix_fwd(logits_arg, NT, NP, edges, slot_t, u_prev.detach(), NA, W, True, sm_escapes=True,
                                     logit_offset=offset)
        torch.testing.assert_close(out, out_ref, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(sm, sm_ref, rtol=1e-5, atol=1e-6)
        d_lg_buf = None
        if strided:
            d_wide = torch.full((m, NT * NP + 7), 7.0, device=DEV)
            d_lg_buf = d_wide[:, 3:3 + NT * NP]
        d_lg, d_u = fused.gate_mix_bwd(logits_arg, NT, NP, edges, slot_t, slot_tower, u_prev.detach(), d_out.contiguous(),
                                       logit_offset=offset, d_logits=d_lg_buf)
        torch.testing.assert_close(d_u, u_prev.grad, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(d_lg.reshape(m, NT, NP), lg_ref.grad, rtol=1e-4, atol=1e-5)
        if strided:
            assert bool((d_wide[:, :3] == 7.0).all()) and bool((d_wide[:, 3 + NT * NP:] == 7.0).all()), "wrote outside"
        again = fused.gate_mix_bwd(logits_arg, NT, NP, edges, slot_t, slot_tower, u_prev.detach(), d_out.contiguous(),
                                   logit_offset=offset)
        assert torch.equal(again[1], d_u), "bit-reproducible"
