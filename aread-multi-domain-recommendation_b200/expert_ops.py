"""The MMoE expert stack + gate mixture as one autograd node on the C-ABI kernels.

    X [B, E] --(4 experts: (Linear -> BatchNorm1d -> ReLU -> Dropout) x 3)--> H [B, 4, h]
    T0[b, g, :] = sum_e gate[b, g, e] * H[b, e, :]                       (model/aread.py:150-153)

Forward per layer: tcgen05 grouped Linear (bf16 operands, fp32 accumulate) -> BN statistics ->
fused BN/ReLU/dropout writing the next layer's bf16 operand.  Backward per layer: BN/ReLU/dropout
gradient (two passes: column sums, then dz) -> tcgen05 weight gradient (split over the samples) and
tcgen05 data gradient.  PyTorch only allocates tensors and orders the launches.
"""
import torch

from . import dense_kernels as dk


class ExpertLayer:
    """Packed parameters of one expert layer (all experts side by side)."""

    def __init__(self, packs, linears, norms, salt):
        self.linears, self.norms = list(linears), list(norms)
        self.weight = packs.add(lambda: [m.weight for m in self.linears])            # flat [G, N, K]
        self.bias = packs.add(lambda: [m.bias for m in self.linears])                # flat [G, N]
        self.gamma = packs.add(lambda: [m.weight for m in self.norms])
        self.beta = packs.add(lambda: [m.bias for m in self.norms])
        self.running_mean = packs.add(lambda: [m.running_mean for m in self.norms])
        self.running_var = packs.add(lambda: [m.running_var for m in self.norms])
        self.salt = salt

    @property
    def tracked(self):                      # num_batches_tracked scalars (buffer objects change on .to())
        return [m.num_batches_tracked for m in self.norms]

    @property
    def params(self):                       # order of the autograd inputs / returned gradients
        return ([m.weight for m in self.linears] + [m.bias for m in self.linears] +
                [m.weight for m in self.norms] + [m.bias for m in self.norms])

    @property
    def groups(self):
        return self.weight.groups

    @property
    def n(self):
        return self.weight.shape[0]

    @property
    def k(self):
        return self.weight.shape[1]


def _hi(t):
    return t[0] if isinstance(t, tuple) else t


def _lo(t):
    return t[1] if isinstance(t, tuple) else None


class ExpertStack(torch.autograd.Function):
    """cfg["precise"]: False = bf16 operands (one tensor-core pass); True = split operands hi + lo,
    three passes into the same accumulator (fp32-grade results)."""

    @staticmethod
    def forward(ctx, x, x_bf16, gate, cfg, *params):
        layers, training, p, seed, precise = (cfg["layers"], cfg["training"], cfg["dropout"], cfg["seed"],
                                              cfg["precise"])
        B = x.shape[0]
        bn_skip = B == 1
        G = layers[0].groups
        if x_bf16 is None:
            x_bf16 = dk.split_bf16(x) if precise else x.to(torch.bfloat16)
        a = x_bf16                                         # tensor, or (hi, lo) in precise mode
        saved_a, saved_z, saved_stats, saved_w = [], [], [], []
        for i, L in enumerate(layers):
            w = dk.split_bf16(L.weight.rows()) if precise else L.weight.rows().to(torch.bfloat16)   # [G * n, k]
            z = dk.grouped_linear(_hi(a), _hi(w), L.bias.rows(), L.n, L.k, G, 0 if i == 0 else L.k,
                                  a_lo=_lo(a), w_lo=_lo(w))
            last = i == len(layers) - 1
            res = dk.bn_act_fwd(z, L.gamma.rows(), L.beta.rows(), L.running_mean.rows(), L.running_var.rows(),
                                training, bn_skip, p, seed, L.salt, None if last else torch.bfloat16,
                                want_lo=precise and not last)
            out, stats = (res[0], res[-1]) if not (precise and not last) else ((res[0], res[1]), res[2])
            if training and not bn_skip:
                torch._foreach_add_(L.tracked, 1)
            saved_a.append(a)
            saved_z.append(z)
            saved_stats.append(stats)
            saved_w.append(w)
            a = out
        L = layers[-1]
        n_gate = gate.shape[1]
        t0 = dk.mmoe_mix_fwd(saved_z[-1], saved_stats[-1], gate.contiguous(), G, n_gate, p if training else 0.0, seed,
                             L.salt)
        ctx.cfg = cfg
        ctx.bn_skip = bn_skip
        ctx.n_gate = n_gate
        ctx.saved_a, ctx.saved_w = saved_a, saved_w        # bf16 operands (plain tensors, no graph)
        ctx.save_for_backward(gate, *saved_z, *saved_stats)
        return t0

    @staticmethod
    def backward(ctx, d_t0):
        cfg = ctx.cfg
        layers, training, seed, precise = cfg["layers"], cfg["training"], cfg["seed"], cfg["precise"]
        p = cfg["dropout"] if training else 0.0
        nl = len(layers)
        saved = ctx.saved_tensors
        gate = saved[0]
        zs = saved[1:1 + nl]
        stats = saved[1 + nl:1 + 2 * nl]
        a_in, ws = ctx.saved_a, ctx.saved_w
        G = layers[0].groups
        d_act, d_gate = dk.mmoe_mix_bwd(zs[-1], stats[-1], gate, d_t0.contiguous(), G, ctx.n_gate, p, seed,
                                        layers[-1].salt)
        grads = {}
        d_x = None

        def transposed(w, fn):
            return tuple(fn(t) for t in w) if isinstance(w, tuple) else fn(w)

        for i in range(nl - 1, -1, -1):
            L = layers[i]
            dz, d_gamma, d_beta, d_bias = dk.bn_act_bwd(zs[i], d_act, stats[i], ctx.bn_skip, p, seed, L.salt,
                                                        want_lo=precise)
            d_w = dk.grouped_wgrad(_hi(dz), _hi(a_in[i]), L.n, L.k, G, 0 if i == 0 else L.k,
                                   dz_lo=_lo(dz), a_lo=_lo(a_in[i]))                      # [G * n, k] fp32
            grads[i] = (d_w.view(G, L.n, L.k), d_bias.view(G, L.n), d_gamma.view(G, L.n), d_beta.view(G, L.n))
            if i > 0:       # data gradient per expert: dA[:, g] = dz[:, g] @ W_g
                wt = transposed(ws[i], lambda t: t.view(G, L.n, L.k).transpose(1, 2).reshape(G * L.k, L.n).contiguous())
                d_act = dk.grouped_linear(_hi(dz), _hi(wt), None, L.k, L.n, G, L.n, a_lo=_lo(dz), w_lo=_lo(wt))
            elif ctx.needs_input_grad[0]:  # all experts read the same X: one GEMM over the stacked outputs
                wt = transposed(ws[0], lambda t: t.t().contiguous())                      # [k, G * n]
                d_x = dk.grouped_linear(_hi(dz), _hi(wt), None, L.k, G * L.n, 1, 0, a_lo=_lo(dz), w_lo=_lo(wt))
        out = []
        for i, L in enumerate(layers):
            for tensors in grads[i]:
                out.extend(tensors.unbind(0))
        # order of `params`: per layer weights, biases, gammas, betas (ExpertLayer.params)
        return (d_x, None, d_gate, None, *out)


def expert_stack(x, x_bf16, gate, layers, training, dropout, seed, precise=False):
    cfg = {"layers": layers, "training": training, "dropout": dropout, "seed": seed, "precise": precise}
    params = [p for L in layers for p in L.params]
    return ExpertStack.apply(x, x_bf16, gate, cfg, *params)
