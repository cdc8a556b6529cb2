"""Kernel-time table of one train step (torch.profiler; cheaper than an ncu launch list)."""
import importlib, sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("aread-multi-domain-recommendation_b200")
wl = importlib.import_module("aread-multi-domain-recommendation_b200.workloads").WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "amazon"]()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda:0")
np.random.seed(0); torch.manual_seed(0)
model = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=bench.N_TOWER, n_domain=wl.n_domain,
                  base_model="mmoe", expert_dims=bench.EXPERT_DIMS, tower_dims=bench.TOWER_DIMS, domain_idx=wl.domain_idx,
                  device=dev, dropout=0.2, config=bench.make_config(wl)).to(dev)
model.reset_for_mask_update()
for d in range(wl.n_domain):
    model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=0.7)
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
crit = torch.nn.BCELoss()
x, y, d = wl.batch(B, seed=1)
x, y = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
def step():
    preds = model(x, mode="domain_mask_bagging", domain_i=d)
    tgt = y.squeeze().float()
    loss = sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0] + model.get_regularization_loss(device=dev)
    model.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
