"""Multi-GPU check of the row-sharded table (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded.py

Every rank builds the same model, keeps an unsharded copy as the checker, shards the table, runs one train step on
its own batch and compares: lookups bit for bit, probabilities, and the shard gradient against the rank-average of
the unsharded dense gradients."""
import importlib, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench

pkg = importlib.import_module("aread-multi-domain-recommendation_b200")
sharding = importlib.import_module("aread-multi-domain-recommendation_b200.sharding")
wl = importlib.import_module("aread-multi-domain-recommendation_b200.workloads").WORKLOADS["amazon"]()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)

def build():
    torch.manual_seed(1); np.random.seed(1)
    m = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=bench.N_TOWER,
                  n_domain=wl.n_domain, base_model="mmoe", expert_dims=bench.EXPERT_DIMS, tower_dims=bench.TOWER_DIMS,
                  domain_idx=wl.domain_idx, device=dev, dropout=0.0, config=bench.make_config(wl)).to(dev)
    m.reset_for_mask_update()
    for d in range(wl.n_domain):
        m.domain_mask[d] = m.generate_mask("rand", d, init_active_percent=0.6)
    m.expert_precision = "bf16x3"
    return m.train()

ref, model = build(), build()
shards = model.shard_table()
B = 4096
x, y, d = wl.batch(B, seed=100 + rank)
x, y = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)

def step(m):
    preds = m(x, mode="domain_mask_bagging", domain_i=d)
    tgt = y.squeeze().float()
    loss = sum(torch.nn.functional.binary_cross_entropy(p, tgt) for p in preds.unbind(0)) / preds.shape[0]
    loss = loss + REG * m.get_regularization_loss(device=dev)
    m.zero_grad(); loss.backward()
    return preds.detach()

REG = 0.0
with torch.no_grad():
    e_ref, e_sh = ref.embedding(x), model.embedding(x)
assert torch.equal(e_ref, e_sh), "sharded lookup differs from the single-table lookup"
p_ref, p_sh = step(ref), step(model)
torch.testing.assert_close(p_sh, p_ref, rtol=1e-5, atol=1e-6)
g_full = ref.embedding.embedding_dict.weight.grad.clone()
dist.all_reduce(g_full, op=dist.ReduceOp.AVG)
want = sharding.split_table(g_full, world, rank)
got = model.embedding.embedding_dict.weight.grad
torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-7)
# with the regulariser in the loss: every rank adds the same 2 * l2 * w, so the average keeps it; on the sharded
# model the term follows the shard parameter (BaseModel.shard_table)
REG = 1.0
step(ref); step(model)
g_full = ref.embedding.embedding_dict.weight.grad.clone()
dist.all_reduce(g_full, op=dist.ReduceOp.AVG)
torch.testing.assert_close(model.embedding.embedding_dict.weight.grad, sharding.split_table(g_full, world, rank),
                           rtol=1e-5, atol=1e-7)
REG = 0.0
step(ref); step(model)
# the same step a few more times: from the third call on both models replay their recorded CUDA-graph sequences
# (the sharded one with its fence and reduce-scatter issued around them)
for _ in range(4):
    p_ref, p_sh = step(ref), step(model)
    torch.testing.assert_close(p_sh, p_ref, rtol=1e-5, atol=1e-6)
    g_full = ref.embedding.embedding_dict.weight.grad.clone()
    dist.all_reduce(g_full, op=dist.ReduceOp.AVG)
    torch.testing.assert_close(model.embedding.embedding_dict.weight.grad, sharding.split_table(g_full, world, rank),
                               rtol=1e-5, atol=1e-7)
assert any(e.bwd is not None for e in model._graphs.entries.values()), "sharded model never replayed a graph"
# dense gradients: flat-bucket all-reduce equals per-tensor averaging
dense = [p for n, p in model.named_parameters() if not n.startswith("embedding.")]
expect = []
for n, p in ref.named_parameters():
    if not n.startswith("embedding."):
        g = torch.zeros_like(p) if p.grad is None else p.grad.clone()
        dist.all_reduce(g, op=dist.ReduceOp.AVG); expect.append(g)
sharding.allreduce_dense_grads(dense)
for a, b in zip(dense, expect):
    torch.testing.assert_close(a.grad, b, rtol=1e-5, atol=1e-8)
# merge_shards round trip
parts = [torch.empty_like(model.embedding.embedding_dict.weight.data) for _ in range(world)]
dist.all_gather(parts, model.embedding.embedding_dict.weight.data)
assert torch.equal(sharding.merge_shards(parts, ref.embedding.embedding_dict.weight.shape[0]), ref.embedding.embedding_dict.weight.data)
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print(f"SHARDED OK world={world}")
dist.destroy_process_group()
