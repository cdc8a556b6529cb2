#!/usr/bin/env python
"""AREAD train throughput on synthetic data of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload amazon|aliccp|cloudtheme] [--batch B]

One step = forward(mode='domain_mask_bagging') + mean-over-towers BCE + L2 regulariser +
zero_grad + backward + Adam.step on one single-domain batch (run.py:663-682).  Prints ONE JSON
line (rank 0): samples/s with inputs resident in HBM (`value`), through the module API from
pinned host buffers (`e2e`), the gather kernel's roofline, and the CPU baseline.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TOWER = (3, 6, 12)
EXPERT_DIMS = (256, 128, 64)
TOWER_DIMS = ((64, 32), (32, 16), (16, 8))
LR, WD = 1e-3, 1e-8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="amazon", choices=["amazon", "aliccp", "cloudtheme", "stress"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per step per GPU")
    ap.add_argument("--cpu-batch", type=int, default=4096, help="rows per step of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=30, help="timed steps of the cpu_baseline leg (about 15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--active", type=float, default=0.7, help="HEMP init_active_percent of the per-domain masks")
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--optimizer", default="fused", choices=["torch", "fused"],
                    help="the library's FusedAdam (one launch; same arithmetic as torch.optim.Adam, tests/test_optim_gpu.py) or "
                         "torch.optim.Adam as run.py:830 builds it")
    ap.add_argument("--graphs", default="prerecord", choices=["prerecord", "lazy"],
                    help="CUDA-graph sequences of the fused node: recorded per domain mask during setup "
                         "(AREAD.record_graphs) or lazily after a few eager steps per mask; AREAD_GRAPHS=0 disables them")
    ap.add_argument("--reg", default="fold", choices=["fold", "loss"],
                    help="gradient of the L2 regulariser: folded into FusedAdam (value still part of the loss) or "
                         "through autograd as in run.py:644")
    ap.add_argument("--loss", default="fused", choices=["fused", "torch"],
                    help="bagging BCE through AREAD.bagging_loss (one kernel) or as the trainer's sum of BCELoss calls")
    return ap.parse_args()


def make_config(wl):
    return types.SimpleNamespace(domain_size={wl.name: list(wl.domain_size)}, dataset_name=wl.name, use_dcn=True,
                                 use_atten=True, n_cross_layers=3, mmoe_n_expert=4, atten_embed_dim=64,
                                 att_head_num=2, att_layer_num=3, att_res=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(wl, batch, steps, warmup, active, seed, dropout=0.2, threads=None):
    """The oracle port (oracle/aread_torch.py: the reference's torch ops restated functionally) timed on
    the host cores: same step definition, fp32, all threads."""
    from oracle import aread_torch as O
    from oracle import synth
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    mh = wl.multi_hot_dict
    spec = O.Spec(one_hot_field_dims=list(wl.one_hot_field_dims), embed_dim=wl.embed_dim,
                  multi_hot_flag=mh["multi_hot_flag"], itemid_idx=wl.itemid_idx, seq_maxlen=wl.seq_maxlen,
                  method=wl.method, n_tower=N_TOWER, n_domain=wl.n_domain, expert_dims=EXPERT_DIMS,
                  tower_dims=TOWER_DIMS, domain_idx=wl.domain_idx, dropout=dropout)
    sd = O.make_leaf_params(synth.deterministic_state(spec))
    opt = O.make_adam(sd, lr=LR, wd=WD)
    np.random.seed(seed)
    hemp = importlib.import_module("aread-multi-domain-recommendation_b200.hemp")
    masks = {}
    times = []
    for step in range(warmup + steps):
        x, y, d = wl.batch(batch, seed=seed + step)
        if d not in masks:
            while True:
                m = hemp.validate_arrays(hemp.full_mask(N_TOWER, active), N_TOWER)
                if m[-1].any():
                    break
            masks[d] = [torch.from_numpy(a) for a in m]
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        t0 = time.perf_counter()
        O.train_step(sd, spec, xt, yt, masks[d], opt, masks="rng")
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return {"value": batch * len(times) / total, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} train steps of {batch} rows ({wl.name}, fp32, dropout {dropout}, {warmup} "
                      f"warm-up; the reference's dead attention branch is not executed), {total:.1f} s"}, total / len(times)


def run_reference(args, wl, rank):
    if rank != 0:
        return
    res, sec_per_step = cpu_reference_rate(wl, args.cpu_batch, max(1, args.steps), max(1, args.warmup), args.active,
                                           args.seed, args.dropout)
    line = {"impl": "reference", "metric": "aread_train_samples_per_sec", "value": res["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(1, args.warmup),
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl.name}_singledomain_B{args.cpu_batch}", "batch_per_step": args.cpu_batch,
                       "n_tower": list(N_TOWER), "embed_dim": wl.embed_dim, "mask_active_percent": args.active,
                       "note": "reference path = torch CPU ops restated in oracle/aread_torch.py (the Python "
                               "reference itself cannot travel to the GPU box); bounded sample of the workload"},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of one gather_kernel launch at the default workload (amazon, B=65536),
# from the committed ncu --set full capture
GATHER_DRAM_BYTES = 21_779_200 + 57_837_824


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, wl, rank, world, local_rank):
    import torch.distributed as dist
    pkg = importlib.import_module("aread-multi-domain-recommendation_b200")
    lib = importlib.import_module("aread-multi-domain-recommendation_b200._lib")
    ops = importlib.import_module("aread-multi-domain-recommendation_b200.embedding_ops")
    lib.load()                                     # fail loudly when the CUDA library is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed + rank)
    np.random.seed(args.seed)                      # same masks on every rank

    model = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=N_TOWER,
                      n_domain=wl.n_domain, base_model="mmoe", expert_dims=EXPERT_DIMS, tower_dims=TOWER_DIMS,
                      domain_idx=wl.domain_idx, device=dev, dropout=args.dropout, config=make_config(wl)).to(dev)
    model.reset_for_mask_update()
    for d in range(wl.n_domain):
        model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=args.active)
    shards = None
    if world > 1:
        # replicas of the dense part, ONE copy of the table: row r lives on rank r % world and is read by its
        # peers over NVLink inside the lookup kernel (sharding.py)
        sharding = importlib.import_module("aread-multi-domain-recommendation_b200.sharding")
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
        shards = model.shard_table()
    model.train()
    fused_adam = importlib.import_module("aread-multi-domain-recommendation_b200.optim").FusedAdam
    adam_cls = fused_adam if args.optimizer == "fused" else torch.optim.Adam
    opt = adam_cls(model.parameters(), lr=LR, betas=(0.9, 0.99), eps=1e-8, weight_decay=WD)
    if args.reg == "fold":                         # L2 gradient applied inside the optimizer step
        if args.optimizer != "fused":
            raise SystemExit("--reg fold needs --optimizer fused")
        model.fold_regularization_into(opt)
    crit = torch.nn.BCELoss()
    table_param = model.embedding.embedding_dict.weight
    dense_params = [p for p in model.parameters() if p is not table_param]

    B = args.batch
    n_batches = args.warmup + args.steps
    host = [wl.batch(B, seed=args.seed + 1000 * rank + i) for i in range(n_batches)]
    host_x = [torch.from_numpy(x).pin_memory() for x, _, _ in host]
    host_y = [torch.from_numpy(y).pin_memory() for _, y, _ in host]
    domains = [d for _, _, d in host]

    def step(x, y, d):
        preds = model(x, mode="domain_mask_bagging", domain_i=d)
        if args.loss == "fused":                   # run.py:672-677 as one kernel (loss_ops.py)
            loss = model.bagging_loss(preds, y)
        else:
            tgt = y.squeeze().float()
            loss = sum(crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0]
        loss = loss + model.get_regularization_loss(device=dev)
        model.zero_grad()
        loss.backward()
        if world > 1:                              # the table gradient arrives reduce-scattered from the backward
            sharding.allreduce_dense_grads(dense_params)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(run_one, first, count):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + count):
            run_one(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident pass
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    if args.graphs == "prerecord":
        # setup, like building the model: the per-mask CUDA-graph launch sequences are recorded once per domain
        # (what a trainer does after every HEMP regroup); parameters, buffers and RNG are left untouched
        # (sharded table: every rank records ALL domains so that the lookup fences / gradient reduce-scatters, which
        # stay outside the recorded sequences, are issued the same number of times everywhere)
        model.record_graphs(dev_x[0], domains=sorted(set(domains)) if world == 1 else range(wl.n_domain))
    for i in range(args.warmup):
        step(dev_x[i], dev_y[i], domains[i])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count()
    ms_dev = timed(lambda i: step(dev_x[i], dev_y[i], domains[i]), args.warmup, args.steps)
    launches = lib.launch_count() - launches0

    # ---- end to end through the module API from pinned host memory (H2D of ids+labels, D2H of the loss)
    # The loss of every step is copied to pinned host memory and READ one step late (after the next step has been
    # enqueued), the way a trainer logs it, so the device->host read does not drain the pipeline; the last step's
    # loss is read inside the timed region as well.
    loss_host = torch.zeros(n_batches, dtype=torch.float32).pin_memory()
    loss_done = [torch.cuda.Event() for _ in range(n_batches)]
    losses = []

    def read_loss(i):
        loss_done[i].synchronize()
        losses.append(float(loss_host[i]))

    def e2e_step(i, first):
        x = host_x[i].to(dev, non_blocking=True)
        y = host_y[i].to(dev, non_blocking=True)
        loss_host[i:i + 1].copy_(step(x, y, domains[i]).detach().reshape(1), non_blocking=True)
        loss_done[i].record()
        if i > first:
            read_loss(i - 1)

    def e2e_run(first, count):
        def run_one(i):
            e2e_step(i, first)
            if i == first + count - 1:
                read_loss(i)
        return run_one
    for i in range(min(2, args.warmup)):
        e2e_run(i, 1)(i)
    ms_e2e = timed(e2e_run(args.warmup, args.steps), args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None         # sampled across both timed regions (all of it under load)
    assert len(losses) == min(2, args.warmup) + args.steps and all(np.isfinite(losses))

    # ---- gather kernel alone, on its launch stream, over the same batches (roofline numerator)
    plan = model.embedding.plan(dev)
    table = model.embedding.embedding_dict.weight.detach()
    for i in range(args.warmup):
        ops.gather(plan, table, dev_x[i])
    stream = torch.cuda.current_stream(dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    g0.record(stream)
    for i in range(args.warmup, args.warmup + args.steps):
        ops.gather(plan, table, dev_x[i])
    g1.record(stream)
    torch.cuda.synchronize(dev)
    gather_ms = g0.elapsed_time(g1) / args.steps

    # ---- the rest of BASELINE's metric, reported beside the headline: eval samples/s (mode 'domain_with_mask',
    # eval(), no_grad -- run.py:712-727) and the scatter (lookup gradient) kernels against HBM
    model.eval()
    if args.graphs == "prerecord":
        with torch.no_grad():
            model.record_graphs(dev_x[0], domains=sorted(set(domains)) if world == 1 else range(wl.n_domain),
                                mode="domain_with_mask", backward=False)

    def eval_step(i):
        with torch.no_grad():
            model(dev_x[i], mode="domain_with_mask", domain_i=domains[i])
    for i in range(args.warmup):
        eval_step(i)
    ms_eval = timed(eval_step, args.warmup, args.steps)
    model.train()
    scatter_ms = scatter_bytes = None
    if world == 1:
        d_out = torch.randn(B, plan.n_fields, plan.embed_dim, device=dev)
        offs = np.asarray(model.embedding.offsets, dtype=np.int64)
        uniq = [int(np.unique(host[i][0].astype(np.int64) + offs[None, :]).size)
                for i in range(args.warmup, args.warmup + args.steps)]
        for i in range(args.warmup):
            ops.scatter(plan, dev_x[i], d_out)
        torch.cuda.synchronize(dev)
        g0.record(stream)
        for i in range(args.warmup, args.warmup + args.steps):
            ops.scatter(plan, dev_x[i], d_out)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        scatter_ms = g0.elapsed_time(g1) / args.steps
        scatter_bytes = float(np.mean([wl.scatter_bytes(B * wl.n_cols, u) for u in uniq]))

    # ---- the tensor-core kernel with the largest share of the step: expert layer 1, [B, E] x [E, 4 * 256] (bf16 in,
    # fp32 accumulate / out), timed alone on its launch stream
    gemm_ms = gemm_flops = None
    if world == 1:
        dk = importlib.import_module("aread-multi-domain-recommendation_b200.dense_kernels")
        n_exp, n1, E = len(model.mmoe_experts), EXPERT_DIMS[0], model.embed_output_dim
        a_op = torch.randn(B, E, device=dev).to(torch.bfloat16)
        w_op = torch.randn(n_exp * n1, E, device=dev).to(torch.bfloat16)
        bias = torch.zeros(n_exp * n1, device=dev)
        out = torch.empty(B, n_exp * n1, device=dev)
        for _ in range(args.warmup):
            dk.grouped_linear(a_op, w_op, bias, n1, E, n_exp, 0, out=out)
        torch.cuda.synchronize(dev)
        g0.record(stream)
        for _ in range(args.steps):
            dk.grouped_linear(a_op, w_op, bias, n1, E, n_exp, 0, out=out)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        gemm_ms = g0.elapsed_time(g1) / args.steps
        gemm_flops = 2.0 * B * n_exp * n1 * E

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    gather_bytes = wl.gather_bytes_per_sample() * B
    achieved = gather_bytes / (gather_ms * 1e-3) / 1e9
    total_samples = B * args.steps * world
    line = {
        "metric": "aread_train_samples_per_sec", "value": total_samples / (ms_dev * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if model.expert_precision == "bf16" else "bf16x3", "data": "synthetic",
        "config": {"workload": f"{wl.name}_singledomain_B{B}", "batch_per_gpu": B, "n_tower": list(N_TOWER),
                   "expert_precision": model.expert_precision + " operands, fp32 accumulate; everything else fp32",
                   "embed_dim": wl.embed_dim, "table_rows": wl.n_rows, "n_cols": wl.n_cols,
                   "mask_active_percent": args.active, "dropout": args.dropout,
                   "optimizer": "torch.optim.Adam" if args.optimizer == "torch" else "aread_b200 FusedAdam",
                   "loss": "AREAD.bagging_loss" if args.loss == "fused" else "sum of torch BCELoss",
                   "l2_regulariser": "value in the loss, gradient folded into FusedAdam" if args.reg == "fold"
                   else "autograd node",
                   "cuda_graphs": ("off (AREAD_GRAPHS=0)" if os.environ.get("AREAD_GRAPHS", "1") == "0" else
                                   "per-mask forward/backward sequences, " + args.graphs),
                   "parallelism": f"dp{world}" + ("" if world == 1 else f" + table row-sharded over {world} GPUs (P2P lookup, "
                                                   "reduce-scatter of the table gradient, flat all-reduce of the rest)"),
                   "l2": "inputs larger than L2: table %d MB + per-step activations" % (wl.n_rows * wl.embed_dim * 4 >> 20)},
        "e2e": {"value": total_samples / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int(host_x[0].numel() * 4 + host_y[0].numel() * 2), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "gather_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": GATHER_DRAM_BYTES if (world == 1 and args.workload == "amazon" and B == 65536) else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/r1_gather_kernel_full.txt); below the algorithmic bytes because the "
                                       "hot rows of the Zipf ids and part of the output stay in the 126 MB L2",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": gather_bytes, "launch_ms": gather_ms},
    }
    line["extra"] = {"eval_samples_per_sec": total_samples / (ms_eval * 1e-3),
                     "eval_mode": "eval(), no_grad, mode='domain_with_mask'"}
    if scatter_ms is not None:
        line["extra"]["scatter"] = {
            "kernels": "scatter_keys + radix sort + scatter_tile + scatter_level (whole aread_scatter_bwd call)",
            "achieved": scatter_bytes / (scatter_ms * 1e-3) / 1e9, "unit": "GB/s", "peak": peaks["hbm_gbs"],
            "frac": scatter_bytes / (scatter_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "launch_ms": scatter_ms,
            "algorithmic_bytes_per_call": scatter_bytes,
            "note": "algorithmic bytes credit lookups*(4+D*4) + unique_rows*D*4 only; the dense [R, D] zero fill the "
                    "reference semantics require (dense gradient) and the sort traffic are not credited"}
    if gemm_ms is not None:
        tf = gemm_flops / (gemm_ms * 1e-3) / 1e12
        out_bytes = B * n_exp * n1 * 4 + B * E * 2
        line["extra"]["grouped_linear_expert_layer1"] = {
            "kernel": "grouped_linear_kernel<128> (tcgen05, TMA in / TMA out)", "launch_ms": gemm_ms,
            "achieved_tflops": tf, "peak_tflops": peaks.get("bf16_tflops"),
            "frac_of_tensor_peak": tf / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
            "achieved_gbs": out_bytes / (gemm_ms * 1e-3) / 1e9, "frac_of_hbm_peak": out_bytes / (gemm_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "note": "K = E = 288 is short: per output element 576 flop against 4 B written, so the fp32 output "
                    "stream (HBM) bounds this GEMM, not the tensor pipe"}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"], _ = cpu_reference_rate(wl, args.cpu_batch, args.cpu_steps, 1, args.active, args.seed,
                                                       args.dropout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    wl = importlib.import_module("aread-multi-domain-recommendation_b200.workloads").WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
