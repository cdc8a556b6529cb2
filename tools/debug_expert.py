import importlib, sys
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from oracle import aread_torch as O, synth
from tests._models import build_model
from tests._util import load_golden
eo = importlib.import_module("aread-multi-domain-recommendation_b200.expert_ops")
DEV='cuda:0'
for name in ('ali_small', 'tiny'):
    fx = load_golden(name); spec = O.Spec(**fx['spec'])
    model = build_model(spec, DEV, dropout=0.0).train()
    x, y = synth.random_batch(spec, fx['B'], seed=11, domain=fx['domain'], pad_id=fx['pad_id'])
    emb, xb = model.embedding.lookup(x.to(DEV), want_bf16=True)
    X = emb.flatten(1)
    gate = torch.softmax(torch.randn(X.shape[0], spec.n_tower[0], spec.n_expert, device=DEV), dim=2)
    T0 = eo.expert_stack(X, xb, gate, model._expert_layers, True, 0.0, 0)
    def ref(round_bf16):
        hs = []
        for e in model.mmoe_experts:
            h = X
            for i in range(len(spec.expert_dims)):
                lin, bn = e.layers[4*i], e.layers[4*i+1]
                if round_bf16:
                    h2 = h.to(torch.bfloat16).float() @ lin.weight.to(torch.bfloat16).float().t() + lin.bias
                else:
                    h2 = h @ lin.weight.t() + lin.bias
                h = F.relu(F.batch_norm(h2, None, None, bn.weight, bn.bias, True, 0.1, 1e-5))
            hs.append(h)
        H = torch.stack(hs, 1)
        return torch.einsum('bge,bew->bgw', gate, H)
    r32, r16 = ref(False), ref(True)
    print(name, 'vs fp32 ref', float((T0-r32).abs().max()), 'vs bf16-emulated ref', float((T0-r16).abs().max()), 'scale', float(r32.abs().max()))
