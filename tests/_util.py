"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("ali_small", "amz_small", "tiny")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def assert_close(actual, expected, rtol, atol, what=""):
    actual = actual.detach().cpu().float()
    expected = expected.detach().cpu().float()
    assert tuple(actual.shape) == tuple(expected.shape), f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    err = (actual - expected).abs()
    tol = atol + rtol * expected.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance (rtol={rtol}, atol={atol}); "
                             f"worst |d|={float(err.reshape(-1)[i]):.3e} at {i}: "
                             f"{float(actual.reshape(-1)[i]):.6e} vs {float(expected.reshape(-1)[i]):.6e}")


def assert_compact(actual, comp, rtol, atol, what=""):
    """Compare a tensor with the compact form written by tests/golden/make_golden.py."""
    actual = actual.detach().cpu().float()
    if "full" in comp:
        return assert_close(actual, comp["full"], rtol, atol, what)
    assert tuple(actual.shape) == tuple(comp["shape"]), what
    flat = actual.reshape(-1)
    assert_close(flat[::comp["stride"]][:comp["sample"].numel()], comp["sample"], rtol, atol, what + " (sample)")
    d = flat.double()
    n = flat.numel()
    sumsq = float((d * d).sum())
    # sums accumulate n roundings; bound by the rms magnitude
    rms = (comp["sumsq"] / n) ** 0.5
    assert abs(float(d.sum()) - comp["sum"]) <= (atol + rtol * rms) * n ** 0.5 * 4 + 1e-12, what + " (sum)"
    assert abs(sumsq - comp["sumsq"]) <= 4 * rtol * comp["sumsq"] + atol * atol * n + 1e-12, what + " (sumsq)"


def assert_after_adam(actual, comp, steps, lr, what="", frac=0.02):
    """Post-Adam weights: Adam normalises each element's step to about lr whatever the size of
    its gradient, so elements whose true gradient is ~0 move by round-off-driven +-lr.  Hard
    bound every element by the largest possible divergence (2*steps*lr) and require all but
    `frac` of them to agree to 5 % of the distance travelled."""
    actual = actual.detach().cpu().float()
    if "full" in comp:
        a, e = actual.reshape(-1), comp["full"].float().reshape(-1)
    else:
        a, e = actual.reshape(-1)[::comp["stride"]][:comp["sample"].numel()], comp["sample"].float()
    err = (a - e).abs()
    hard = 2.0 * steps * lr + 1e-4 * e.abs()
    assert bool((err <= hard).all()), f"{what}: max |d| {float(err.max()):.3e} beyond 2*steps*lr"
    soft = 0.05 * steps * lr + 1e-4 * e.abs()
    n_bad = int((err > soft).sum())
    assert n_bad <= max(4, int(frac * err.numel())), f"{what}: {n_bad}/{err.numel()} beyond 5% of the Adam travel"


import re as _re

GRAD_FAMILIES = [
    ("table", _re.compile(r"^embedding\.")),
    ("linear_cross", _re.compile(r"^(linear|cn)\.")),
    ("expert_weights", _re.compile(r"^mmoe_experts\.\d+\.layers\.(0|4|8)\.weight$")),
    ("expert_bn", _re.compile(r"^mmoe_experts\.\d+\.layers\.(1|5|9)\.")),
    ("mmoe_gates", _re.compile(r"^mmoe_gates\.")),
    ("tower_weights", _re.compile(r"^towers\.\d+\.\d+\.layers\.(0|4)\.weight$")),
    ("tower_bn", _re.compile(r"^towers\.\d+\.\d+\.layers\.(1|5)\.")),
    ("tower_gates", _re.compile(r"^(tower_gates|group_embedding)\.")),
    ("heads", _re.compile(r"^towers_linear\.")),
]
PRE_BN_BIAS = _re.compile(r"\.layers\.(0|4|8)\.bias$")      # true gradient is zero (BatchNorm removes the bias)


def family_errors(pairs):
    """pairs: iterable of (state_dict key, gradient, reference gradient) as flat CPU tensors of equal length.
    -> {family: (normalised error ||g - ref|| / ||ref||, 1 - cosine)} over the concatenation of the family's tensors.
    Single small tensors behind the towers are residues of cancelling terms (their own normalised error is dominated
    by round-off amplification); the family vector is what an optimizer step sees."""
    buckets = {}
    for k, g, ref in pairs:
        if PRE_BN_BIAS.search(k):
            continue
        for fam, pat in GRAD_FAMILIES:
            if pat.search(k):
                a, b = buckets.setdefault(fam, ([], []))
                a.append(g.reshape(-1).double())
                b.append(ref.reshape(-1).double())
                break
    out = {}
    for fam, (a, b) in buckets.items():
        a, b = torch.cat(a), torch.cat(b)
        nb = float(b.norm())
        if nb > 0:
            out[fam] = (float((a - b).norm()) / nb, 1.0 - float(a @ b) / (float(a.norm()) * nb + 1e-300))
    return out
