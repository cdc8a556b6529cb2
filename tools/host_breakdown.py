"""Host-side (enqueue) time of each phase of a train step, B=1024 so the GPU is never the limit."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
pkg = importlib.import_module("aread-multi-domain-recommendation_b200")
wl = importlib.import_module("aread-multi-domain-recommendation_b200.workloads").WORKLOADS["amazon"]()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(record):
    t = time.perf_counter(); preds = model(x, mode="domain_mask_bagging", domain_i=d)
    if record: tick("forward", t)
    t = time.perf_counter(); tgt = y.squeeze().float(); loss = sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0]
    if record: tick("bce", t)
    t = time.perf_counter(); loss = loss + model.get_regularization_loss(device=dev)
    if record: tick("reg", t)
    t = time.perf_counter(); model.zero_grad()
    if record: tick("zero_grad", t)
    t = time.perf_counter(); loss.backward()
    if record: tick("backward", t)
    t = time.perf_counter(); opt.step()
    if record: tick("opt.step", t)
for _ in range(5): step(False)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N): step(True)
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / N
for k, v in T.items(): print(f"{k:10s} {1e3 * v / N:7.3f} ms")
print(f"total      {1e3 * total:7.3f} ms/step")
if len(sys.argv) > 2:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10): step(False)
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
