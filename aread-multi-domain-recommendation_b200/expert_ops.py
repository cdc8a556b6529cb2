"""Packed views of one expert / tower layer family: the Linear and BatchNorm1d parameters of all experts (or all
towers of a level) side by side, in the order the fused node's backward returns their gradients (fused.py), plus
the bf16 hi / lo operand helpers of the split-precision GEMMs."""


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
