import importlib, sys, re
sys.path.insert(0, '/root/repo')
import torch
from oracle import aread_torch as O, synth
from tests._models import build_model
from tests._util import load_golden
DEV='cuda:0'
for name in ('ali_small','amz_small','tiny'):
  for mk in ('full','sparse'):
    fx = load_golden(name); spec = O.Spec(**fx['spec'])
    x, y = synth.random_batch(spec, fx['B'], seed=11, domain=fx['domain'], pad_id=fx['pad_id'])
    mask = fx['masks'][mk]
    res = {}
    for tag, dt in (('fp32', None), ('bf16', torch.bfloat16)):
        sp = O.Spec(**fx['spec'], expert_operand_dtype=dt)
        sd = O.make_leaf_params(synth.deterministic_state(sp))
        out = O.forward(sd, sp, x, 'domain_mask_bagging', mask, training=True)
        (O.bagging_loss(out['y'], y) + O.reg_loss(sd, sp)).backward()
        res[tag] = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    model = build_model(spec, DEV, dropout=0.0).train()
    preds = model(x.to(DEV), mode='domain_mask_bagging', current_mask=[m.to(DEV) for m in mask])
    tgt = y.to(DEV).squeeze().float()
    loss = sum(torch.nn.functional.binary_cross_entropy(p, tgt) for p in preds.unbind(0))/preds.shape[0] + model.get_regularization_loss(device=torch.device(DEV))
    model.zero_grad(); loss.backward()
    worst = {}
    for k, p in model.named_parameters():
        if p.grad is None or k not in res['fp32']: continue
        fam = re.sub(r'\.\d+', '', k)
        g = p.grad.cpu()
        e32 = float((g-res['fp32'][k]).norm()/(res['fp32'][k].norm()+1e-30)); e16 = float((g-res['bf16'][k]).norm()/(res['bf16'][k].norm()+1e-30))
        o = float((res['bf16'][k]-res['fp32'][k]).norm()/(res['fp32'][k].norm()+1e-30))
        w = worst.setdefault(fam, [0,0,0]); w[0]=max(w[0],e32); w[1]=max(w[1],e16); w[2]=max(w[2],o)
    print(name, mk)
    for fam, (a,b,c) in worst.items():
        if 'layers.bias' in fam and ('layers.0' in fam): pass
        print(f'   {fam:40s} gpu-vs-fp32 {a:.2e}  gpu-vs-bf16oracle {b:.2e}  bf16oracle-vs-fp32 {c:.2e}')
