"""Times the HEI tower-layer entry points (aread_hei_layer_fwd / _bwd) at the shapes of one AliCCP-shaped step.
The path (tensor cores, csrc/hei_tc.cu, or CUDA cores, csrc/hei.cu) is fixed per process by AREAD_HEI_TC /
AREAD_HEI_TC_BWD; run it once per setting.  Prints one JSON line per shape: us per launch and the HBM rate of the
algorithmic bytes (input + output once forward; z + d_out + input + d_in once backward)."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ho = importlib.import_module("aread-multi-domain-recommendation_b200.hei_ops")

DEV = "cuda:0"
M = int(os.environ.get("HEI_ROWS", 65536))
SHAPES = [(3, 64, 64), (3, 64, 32), (4, 32, 32), (4, 32, 16), (8, 16, 16), (8, 16, 8), (6, 32, 32), (12, 16, 16)]
if os.environ.get("HEI_SHAPES"):
    SHAPES = [tuple(int(v) for v in s.split(",")) for s in os.environ["HEI_SHAPES"].split(";")]
N_TIMED = int(os.environ.get("HEI_ITERS", 20))


def timed(fn, n=N_TIMED):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    total = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / n * 1e3


def main():
    gen = torch.Generator(device=DEV).manual_seed(0)
    for G, K, N in SHAPES:
        rnd = lambda *s: torch.randn(*s, device=DEV, generator=gen)
        zp, w, b = rnd(M, G * K), 0.3 * rnd(G, N, K), rnd(G, N)
        gamma, beta = 1 + 0.1 * rnd(G * N), 0.1 * rnd(G * N)
        rm, rv = torch.zeros(G * N, device=DEV), torch.ones(G * N, device=DEV)
        sp = torch.stack([zp.mean(0), 1 / zp.std(0), 1 / zp.std(0), -zp.mean(0) / zp.std(0)]).contiguous()
        d_out = rnd(M, G * N)
        for src_bn in (False, True):
            saved = sp if src_bn else None
            z, s = ho.layer_fwd(zp, saved, 7, w, b, gamma, beta, rm, rv, G, K, N, True, False, 0.2, 11)
            coef, _ = ho.bn_bwd_coef(z, d_out, s, False, 0.2, 11, 9)
            f = timed(lambda: ho.layer_fwd(zp, saved, 7, w, b, gamma, beta, rm, rv, G, K, N, True, False, 0.2, 11))
            g = timed(lambda: ho.layer_bwd(z, d_out, s, coef, 0.2, 9, 11, False, zp, saved, 7, w, G, K, N))
            fb, bb = 4 * M * G * (K + N), 4 * M * G * (2 * N + 2 * K)
            print(json.dumps({"groups": G, "k": K, "n": N, "src_bn": src_bn, "rows": M,
                              "tc": os.environ.get("AREAD_HEI_TC", "1"), "tc_bwd": os.environ.get("AREAD_HEI_TC_BWD", "1"),
                              "fwd_us": round(f, 1), "fwd_gbs": round(fb / f / 1e3, 1),
                              "bwd_us": round(g, 1), "bwd_gbs": round(bb / g / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
