"""Per-phase time (host enqueue + device, synchronised after every phase) of the N-GPU train step.
Run under torchrun; rank 0 prints."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
pkg = importlib.import_module("aread-multi-domain-recommendation_b200")
sharding = importlib.import_module("aread-multi-domain-recommendation_b200.sharding")
wl = importlib.import_module("aread-multi-domain-recommendation_b200.workloads").WORKLOADS["amazon"]()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
np.random.seed(0); torch.manual_seed(rank)
model = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=bench.N_TOWER, n_domain=wl.n_domain,
                  base_model="mmoe", expert_dims=bench.EXPERT_DIMS, tower_dims=bench.TOWER_DIMS, domain_idx=wl.domain_idx,
                  device=dev, dropout=0.2, config=bench.make_config(wl)).to(dev)
model.reset_for_mask_update()
for d in range(wl.n_domain):
    model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=0.7)
if world > 1:
    for p in model.parameters():
        dist.broadcast(p.data, src=0)
    model.shard_table()
model.train()
Adam = importlib.import_module("aread-multi-domain-recommendation_b200.optim").FusedAdam if os.environ.get("OPT", "fused") == "fused" else torch.optim.Adam
opt = Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
if os.environ.get("REG", "fold") == "fold" and os.environ.get("OPT", "fused") == "fused":
    model.fold_regularization_into(opt)
crit = torch.nn.BCELoss()
table = model.embedding.embedding_dict.weight
dense = [p for p in model.parameters() if p is not table]
batches = []
for i in range(8):
    x, y, d = wl.batch(B, seed=1 + 100 * rank + i)
    batches.append((torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), d))
T = {}
SYNC = os.environ.get("PHASE_SYNC", "1") == "1"
def tick(name, t0):
    if SYNC: torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(i, record):
    x, y, d = batches[i % len(batches)]
    t = time.perf_counter(); preds = model(x, mode="domain_mask_bagging", domain_i=d)
    if record: tick("forward", t)
    t = time.perf_counter(); loss = model.bagging_loss(preds, y)
    if record: tick("bce", t)
    t = time.perf_counter(); loss = loss + model.get_regularization_loss(device=dev)
    if record: tick("reg", t)
    t = time.perf_counter(); model.zero_grad()
    if record: tick("zero_grad", t)
    t = time.perf_counter(); loss.backward()
    if record: tick("backward", t)
    if world > 1:
        t = time.perf_counter(); sharding.allreduce_dense_grads(dense)
        if record: tick("allreduce", t)
    t = time.perf_counter(); opt.step()
    if record: tick("opt.step", t)
for i in range(int(os.environ.get('WARM', 5))): step(i, False)
torch.cuda.synchronize()
if world > 1: dist.barrier()
N = 16
t0 = time.perf_counter()
for i in range(N): step(i, True)
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / N
if rank == 0:
    print(f"world {world} B {B} sync {SYNC} cpus {os.cpu_count()} threads {torch.get_num_threads()}")
    for k, v in T.items(): print(f"{k:10s} {1e3 * v / N:7.3f} ms")
    print(f"total      {1e3 * total:7.3f} ms/step")
if os.environ.get("KINETO") == "1":
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(4): step(i, False)
        torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
if world > 1:
    dist.destroy_process_group()
