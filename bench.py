#!/usr/bin/env python
"""AREAD train throughput on synthetic data of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload aliccp|amazon|cloudtheme|stress] [--batch B]

One step = forward(mode='domain_mask_bagging') + mean-over-towers BCE + L2 regulariser +
zero_grad + backward + Adam.step on one single-domain batch (run.py:663-682).  Prints ONE JSON
line (rank 0): samples/s with inputs resident in HBM (`value`), through the module API from
pinned host buffers (`e2e`), the roofline of the kernel with the largest share of the step plus the
lookup kernel's, the CPU baseline, and under `extra` the same step at other batch sizes, with the
unedited trainer calls, on torch-eager CUDA ops (the same-box GPU bar), eval throughput, the lookup
gradient and the HEMP regroup loop.

The default workload is the AliCCP-scale configuration (BASELINE.json configs[2]): it is the one the
metric is quoted on and it fits one GPU.
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

PKG = "aread-multi-domain-recommendation_b200"
N_TOWER = (3, 6, 12)
EXPERT_DIMS = (256, 128, 64)
TOWER_DIMS = ((64, 32), (32, 16), (16, 8))
LR, WD = 1e-3, 1e-8

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/), keyed by (kernel, workload, batch); None where no capture exists
NCU_DRAM_BYTES = {
    # profiles/r2_grouped_linear_kernel_full.txt, grouped_linear_kernel<128, STATS> (expert layer 1)
    ("grouped_linear_kernel", "aliccp", 65536): 98_006_016 + 92_533_504,
}
# the lookup as the train step launches it (fp32 rows + the split bf16 copy for the tensor-core products), from
# profiles/r2_gather_kernel_full.txt: 90.5 us, 42.2 MB read + 331.5 MB written
GATHER_IN_STEP = {("aliccp", 65536): {"launch_us_ncu": 90.5, "dram_bytes": 42_155_520 + 331_523_072,
                                      "note": "fp32 [B, F, D] output (193 MB) + hi / lo bf16 copies (2 x 96 MB) written by "
                                              "the same kernel; table rows are served from L2 (Zipf ids, 146 MB table)"}}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="aliccp", choices=["amazon", "aliccp", "cloudtheme", "stress"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per step per GPU")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the cpu_baseline leg (about 15-25 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline and rooflines only (profiling runs)")
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
    ap.add_argument("--sustained-steps", type=int, default=200, help="steps of the extra.sustained leg")
    ap.add_argument("--step-graph", default="on", choices=["on", "off"],
                    help="the whole step (forward + loss + backward + FusedAdam) as one CUDA graph per mask "
                         "(step_graph.GraphedTrainStep; single GPU, fused optimizer / loss / folded regulariser only)")
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
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback"


def workload_config(wl, args, world, model=None):
    """The `config` object of the JSON line; the reference arm prints the same one."""
    cfg = {"workload": f"{wl.name}_singledomain_B{args.batch}", "batch_per_gpu": args.batch, "n_tower": list(N_TOWER),
           "embed_dim": wl.embed_dim, "table_rows": wl.n_rows, "n_cols": wl.n_cols,
           "mask_active_percent": args.active, "dropout": args.dropout}
    return cfg


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


# ----------------------------------------------------------------------------------------- oracle arms
def oracle_rate(wl, batch, steps, warmup, active, seed, dropout=0.2, threads=None, device="cpu"):
    """The oracle port (oracle/aread_torch.py: the reference's torch ops restated functionally) timed as the
    reference path: same step definition, fp32.  device='cpu': all host threads (the reference arm / cpu_baseline);
    device='cuda:N': torch-eager CUDA kernels on the same box (the 'same-box GPU bar' of SURVEY 8d)."""
    from oracle import aread_torch as O
    from oracle import synth
    on_gpu = device != "cpu"
    threads = threads or os.cpu_count()
    if not on_gpu:
        torch.set_num_threads(threads)
    mh = wl.multi_hot_dict
    spec = O.Spec(one_hot_field_dims=list(wl.one_hot_field_dims), embed_dim=wl.embed_dim,
                  multi_hot_flag=mh["multi_hot_flag"], itemid_idx=wl.itemid_idx, seq_maxlen=wl.seq_maxlen,
                  method=wl.method, n_tower=N_TOWER, n_domain=wl.n_domain, expert_dims=EXPERT_DIMS,
                  tower_dims=TOWER_DIMS, domain_idx=wl.domain_idx, dropout=dropout)
    state = synth.deterministic_state(spec)
    if on_gpu:
        state = {k: v.to(device) for k, v in state.items()}
    sd = O.make_leaf_params(state)
    opt = O.make_adam(sd, lr=LR, wd=WD)
    np.random.seed(seed)
    hemp = importlib.import_module(PKG + ".hemp")
    masks = {}
    times = []
    for step in range(warmup + steps):
        x, y, d = wl.batch(batch, seed=seed + step)
        if d not in masks:
            while True:
                m = hemp.validate_arrays(hemp.full_mask(N_TOWER, active), N_TOWER)
                if m[-1].any():
                    break
            masks[d] = [torch.from_numpy(a).to(device) for a in m]
        xt, yt = torch.from_numpy(x).to(device), torch.from_numpy(y).to(device)
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.train_step(sd, spec, xt, yt, masks[d], opt, masks="rng")
        if on_gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    where = "torch-eager CUDA ops on the same GPU" if on_gpu else f"{threads} host threads"
    return {"value": batch * len(times) / total, "unit": "samples/s", "cores": threads if not on_gpu else 0,
            "kind": "port",
            "sample": f"{len(times)} train steps of {batch} rows ({wl.name}, fp32, dropout {dropout}, {warmup} "
                      f"warm-up, {where}; the reference's dead attention branch is not executed), {total:.1f} s"}, \
        total / len(times)


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    # a bounded sample of the workload: the same batch as our arm when the whole run stays within a few minutes
    batch = args.batch if args.batch * (steps + warmup) <= 2_000_000 else 4096
    res, sec_per_step = oracle_rate(wl, batch, steps, warmup, args.active, args.seed, args.dropout)
    cfg = workload_config(wl, args, world)
    cfg["note"] = ("reference path = torch CPU ops restated in oracle/aread_torch.py (the Python reference itself "
                   f"cannot travel to the GPU box); each step is {batch} rows of the workload")
    line = {"impl": "reference", "metric": "aread_train_samples_per_sec", "value": res["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
class Trainer:
    """Model + optimizer + the train step of run.py:663-682 in the chosen variant."""

    def __init__(self, pkg, wl, args, dev, world, optimizer, reg, loss, graphs):
        self.args, self.wl, self.dev, self.world = args, wl, dev, world
        self.loss_kind, self.graphs = loss, graphs
        self.model = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=N_TOWER,
                               n_domain=wl.n_domain, base_model="mmoe", expert_dims=EXPERT_DIMS,
                               tower_dims=TOWER_DIMS, domain_idx=wl.domain_idx, device=dev, dropout=args.dropout,
                               config=make_config(wl)).to(dev)
        model = self.model
        model.reset_for_mask_update()
        for d in range(wl.n_domain):
            model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=args.active)
        self.sharding = None
        if world > 1:
            import torch.distributed as dist
            # replicas of the dense part, ONE copy of the table: row r lives on rank r % world and is read by its
            # peers over NVLink inside the lookup kernel (sharding.py)
            self.sharding = importlib.import_module(PKG + ".sharding")
            for p in model.parameters():
                dist.broadcast(p.data, src=0)
            model.shard_table()
        model.train()
        fused_adam = importlib.import_module(PKG + ".optim").FusedAdam
        adam_cls = fused_adam if optimizer == "fused" else torch.optim.Adam
        self.opt = adam_cls(model.parameters(), lr=LR, betas=(0.9, 0.99), eps=1e-8, weight_decay=WD)
        if reg == "fold":                          # L2 gradient applied inside the optimizer step
            if optimizer != "fused":
                raise SystemExit("--reg fold needs --optimizer fused")
            model.fold_regularization_into(self.opt)
        self.crit = torch.nn.BCELoss()
        table_param = model.embedding.embedding_dict.weight
        self.dense_params = [p for p in model.parameters() if p is not table_param]
        self.runner = None
        if (args.step_graph == "on" and world == 1 and optimizer == "fused" and reg == "fold" and loss == "fused" and
                graphs == "prerecord" and os.environ.get("AREAD_GRAPHS", "1") != "0"):
            self.runner = importlib.import_module(PKG + ".step_graph").GraphedTrainStep(model, self.opt)

    def record(self, x, domains, mode="domain_mask_bagging", backward=True, y=None):
        if self.runner is not None and backward and y is not None:
            # setup: one eager pass (sizes the arena, loads every kernel) and one recording pass per distinct mask
            seen = set()
            for d in sorted(set(domains)):
                serial = self.model.mask_info(self.model.domain_mask[d]).serial
                if serial not in seen:
                    seen.add(serial)
                    for _ in range(3):
                        self.runner(x, y, d)
            return
        if self.graphs == "prerecord":
            # setup, like building the model: the per-mask CUDA-graph launch sequences are recorded once per distinct
            # mask (what a trainer does after every HEMP regroup); parameters, buffers and RNG are left untouched
            # (sharded table: every rank records ALL domains so that the lookup fences / gradient reduce-scatters,
            # which stay outside the recorded sequences, are issued the same number of times everywhere)
            doms = sorted(set(domains)) if self.world == 1 else range(self.wl.n_domain)
            self.model.record_graphs(x, domains=doms, mode=mode, backward=backward)

    def step(self, x, y, d):
        if self.runner is not None:
            return self.runner(x, y, d)
        model = self.model
        preds = model(x, mode="domain_mask_bagging", domain_i=d)
        if self.loss_kind == "fused":              # run.py:672-677 as one kernel (loss_ops.py)
            loss = model.bagging_loss(preds, y)
        else:
            tgt = y.squeeze().float()
            loss = sum(self.crit(p, tgt) for p in preds.unbind(dim=0)) / preds.shape[0]
        loss = loss + model.get_regularization_loss(device=self.dev)
        model.zero_grad()
        loss.backward()
        if self.world > 1:                         # the table gradient arrives reduce-scattered from the backward
            self.sharding.allreduce_dense_grads(self.dense_params)
        self.opt.step()
        return loss


def make_batches(wl, B, n, seed, rank, dev):
    host = [wl.batch(B, seed=seed + 1000 * rank + i) for i in range(n)]
    hx = [torch.from_numpy(x).pin_memory() for x, _, _ in host]
    hy = [torch.from_numpy(y).pin_memory() for _, y, _ in host]
    return host, hx, hy, [d for _, _, d in host]


def run_ours(args, wl, rank, world, local_rank):
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    lib = importlib.import_module(PKG + "._lib")
    ops = importlib.import_module(PKG + ".embedding_ops")
    lib.load()                                     # fail loudly when the CUDA library is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed + rank)
    np.random.seed(args.seed)                      # same masks on every rank

    tr = Trainer(pkg, wl, args, dev, world, args.optimizer, args.reg, args.loss, args.graphs)
    model = tr.model
    B = args.batch
    n_batches = args.warmup + args.steps
    host, host_x, host_y, domains = make_batches(wl, B, n_batches, args.seed, rank, dev)

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
    tr.record(dev_x[0], domains, y=dev_y[0])
    for i in range(args.warmup):
        tr.step(dev_x[i], dev_y[i], domains[i])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count()
    ms_dev = timed(lambda i: tr.step(dev_x[i], dev_y[i], domains[i]), args.warmup, args.steps)
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
        loss_host[i:i + 1].copy_(tr.step(x, y, domains[i]).detach().reshape(1), non_blocking=True)
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

    stream = torch.cuda.current_stream(dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def kernel_ms(fn, n_warm, n):
        """Average duration of `fn`'s launches, timed alone on their launch stream."""
        for i in range(n_warm):
            fn(i)
        torch.cuda.synchronize(dev)
        g0.record(stream)
        for i in range(n_warm, n_warm + n):
            fn(i)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        return g0.elapsed_time(g1) / n

    # ---- lookup kernel alone over the same batches
    plan = model.embedding.plan(dev)
    table = model.embedding.embedding_dict.weight.detach()
    gather_ms = kernel_ms(lambda i: ops.gather(plan, table, dev_x[i % n_batches]), args.warmup, args.steps)

    # ---- the kernel with the largest share of the step (profiles/r2_launches_*.txt): expert layer 1 on the tensor
    # cores, [B, E] x [E, 4 * 256] bf16 in, fp32 accumulate
    gemm = None
    if world == 1:
        gemm = expert_gemm_roofline(model, B, dev, kernel_ms, args)

    extra = {}
    if not args.no_extra:
        extra = extra_legs(args, wl, tr, pkg, lib, ops, dev, rank, world, timed, kernel_ms, host, dev_x, dev_y, domains)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    gather_bytes = wl.gather_bytes_per_sample() * B
    achieved = gather_bytes / (gather_ms * 1e-3) / 1e9
    total_samples = B * args.steps * world
    cfg = workload_config(wl, args, world)
    cfg.update({
        "expert_precision": model.expert_precision + " operands, fp32 accumulate; everything else fp32",
        "optimizer": "torch.optim.Adam" if args.optimizer == "torch" else "aread_b200 FusedAdam",
        "loss": "AREAD.bagging_loss" if args.loss == "fused" else "sum of torch BCELoss",
        "l2_regulariser": "value in the loss, gradient folded into FusedAdam" if args.reg == "fold" else "autograd node",
        "cuda_graphs": ("off (AREAD_GRAPHS=0)" if os.environ.get("AREAD_GRAPHS", "1") == "0" else
                        ("whole train step per mask (step_graph.GraphedTrainStep)" if tr.runner is not None else
                         "per-mask forward/backward sequences, " + args.graphs)),
        "parallelism": f"dp{world}" + ("" if world == 1 else f" + table row-sharded over {world} GPUs (P2P lookup, "
                                       "sparse exchange of the table gradient, flat all-reduce of the rest)"),
        "l2": "inputs larger than L2: table %d MB + per-step activations" % (wl.n_rows * wl.embed_dim * 4 >> 20)})
    gather_roof = {"bound": "hbm", "kernel": "gather_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                   "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                   "traffic": NCU_DRAM_BYTES.get(("gather_kernel", args.workload, B)),
                   "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, profiles/); "
                                   "below the algorithmic bytes when hot rows of the Zipf ids and part of the output "
                                   "stay in the 126 MB L2: `achieved` is then an L2-assisted rate, the DRAM-level "
                                   "rate is traffic / launch_ms",
                   "peak_source": peak_src, "algorithmic_bytes_per_launch": gather_bytes, "launch_ms": gather_ms}
    if gather_roof["traffic"]:
        gather_roof["dram_level_gbs"] = gather_roof["traffic"] / (gather_ms * 1e-3) / 1e9
    in_step = GATHER_IN_STEP.get((args.workload, B))
    if in_step:
        gather_roof["in_step"] = dict(in_step, dram_level_gbs=in_step["dram_bytes"] / (in_step["launch_us_ncu"] * 1e-6) / 1e9,
                                      frac_of_hbm_peak=in_step["dram_bytes"] / (in_step["launch_us_ncu"] * 1e-6) / 1e9
                                      / peaks["hbm_gbs"])
    line = {
        "metric": "aread_train_samples_per_sec", "value": total_samples / (ms_dev * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if model.expert_precision == "bf16" else "bf16x3", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": total_samples / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": int(host_x[0].numel() * 4 + host_y[0].numel() * 2), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if gemm is not None:
        line["roofline"] = gemm
        line["roofline_gather"] = gather_roof
    else:
        line["roofline"] = gather_roof
    line["extra"] = extra
    if not args.no_cpu_baseline and world == 1:
        cpu_b = B if B * (args.cpu_steps + 1) <= 300_000 else 4096
        line["cpu_baseline"], _ = oracle_rate(wl, cpu_b, args.cpu_steps, 1, args.active, args.seed, args.dropout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def expert_gemm_roofline(model, B, dev, kernel_ms, args):
    dk = importlib.import_module(PKG + ".dense_kernels")
    peaks, peak_src = load_peaks()
    n_exp, n1, E = len(model.mmoe_experts), EXPERT_DIMS[0], model.embed_output_dim
    a_op = torch.randn(B, E, device=dev).to(torch.bfloat16)
    w_op = torch.randn(n_exp * n1, E, device=dev).to(torch.bfloat16)
    run, out_bytes, what = dk.bench_expert_layer1(a_op, w_op, n1, E, n_exp)
    ms = kernel_ms(lambda i: run(), max(3, args.warmup), max(10, args.steps))
    flops = 2.0 * B * n_exp * n1 * E
    tf = flops / (ms * 1e-3) / 1e12
    bytes_alg = B * E * 2 + n_exp * n1 * E * 2 + out_bytes
    return {"bound": "tensor", "kernel": what, "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": tf / peaks["bf16_tflops"], "peak_source": peak_src + " (burst: kernel timed alone)",
            "traffic": NCU_DRAM_BYTES.get(("grouped_linear_kernel", args.workload, B)),
            "algorithmic_flops_per_launch": flops, "launch_ms": ms,
            "hbm": {"algorithmic_bytes_per_launch": bytes_alg, "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "note": "expert layer 1 of all 4 experts, [B, E] x [E, 4*256]: 2*B*E*1024 flop per launch "
                    "(SURVEY 8d: 4*E*256 MAC per sample); the operand / output streams are listed under `hbm`"}


def extra_legs(args, wl, tr, pkg, lib, ops, dev, rank, world, timed, kernel_ms, host, dev_x, dev_y, domains):
    """Everything beside the headline: eval rate, lookup gradient, sustained run, other batch sizes, the unedited
    trainer calls, torch-eager on the same GPU, the HEMP regroup loop."""
    model, B = tr.model, args.batch
    peaks, _ = load_peaks()
    n_batches = len(dev_x)
    extra = {}

    # ---- 200 steps back to back (the headline's K comes from the driver)
    n_sus = args.sustained_steps
    if n_sus > 0:
        ms = timed(lambda i: tr.step(dev_x[i % n_batches], dev_y[i % n_batches], domains[i % n_batches]), 0, n_sus)
        extra["sustained"] = {"steps": n_sus, "ms_per_step": ms / n_sus, "samples_per_sec": B * n_sus * world / (ms * 1e-3),
                              "timed_region_s": ms * 1e-3}

    # ---- eval samples/s (mode 'domain_with_mask', eval(), no_grad -- run.py:712-727)
    model.eval()
    with torch.no_grad():
        tr.record(dev_x[0], domains, mode="domain_with_mask", backward=False)

    def eval_step(i):
        with torch.no_grad():
            model(dev_x[i], mode="domain_with_mask", domain_i=domains[i])
    for i in range(args.warmup):
        eval_step(i)
    ms_eval = timed(eval_step, args.warmup, args.steps)
    model.train()
    extra["eval_samples_per_sec"] = B * args.steps * world / (ms_eval * 1e-3)
    extra["eval_mode"] = "eval(), no_grad, mode='domain_with_mask'"
    if world > 1:
        return extra

    # ---- HEI tower layers on the tensor cores (csrc/hei_tc.cu) at the widths of this model, all towers of a level
    extra["hei_layers"] = hei_leg(model, B, dev, kernel_ms, peaks, args)

    # ---- lookup gradient (whole aread_scatter_bwd call) against HBM
    plan = model.embedding.plan(dev)
    d_out = torch.randn(B, plan.n_fields, plan.embed_dim, device=dev)
    offs = np.asarray(model.embedding.offsets, dtype=np.int64)
    uniq = [int(np.unique(host[i][0].astype(np.int64) + offs[None, :]).size) for i in range(n_batches)]
    scatter_ms = kernel_ms(lambda i: ops.scatter(plan, dev_x[i % n_batches], d_out), args.warmup, args.steps)
    scatter_bytes = float(np.mean([wl.scatter_bytes(B * wl.n_cols, u) for u in uniq]))
    extra["scatter"] = {
        "kernels": "whole aread_scatter_bwd call (keys, sort, segmented in-order reduction, dense zero fill)",
        "achieved": scatter_bytes / (scatter_ms * 1e-3) / 1e9, "unit": "GB/s", "peak": peaks["hbm_gbs"],
        "frac": scatter_bytes / (scatter_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "launch_ms": scatter_ms,
        "algorithmic_bytes_per_call": scatter_bytes,
        "note": "algorithmic bytes credit lookups*(4+D*4) + unique_rows*D*4 only; the dense [R, D] zero fill the "
                "reference semantics require (dense gradient) and the sort traffic are not credited"}

    # ---- the reference's own batch sizes (main.py:22 bs=1024; SURVEY 8d asks 1024 / 8192 / 65536)
    for b_small in (1024, 8192):
        if b_small >= B:
            continue
        n = 40
        _, hx, hy, doms = make_batches(wl, b_small, n, args.seed + 77, rank, dev)
        dx, dy = [t.to(dev) for t in hx], [t.to(dev) for t in hy]
        tr.record(dx[0], doms, y=dy[0])
        for i in range(10):
            tr.step(dx[i], dy[i], doms[i])
        ms = timed(lambda i: tr.step(dx[i], dy[i], doms[i]), 10, n - 10)
        extra[f"batch_{b_small}"] = {"ms_per_step": ms / (n - 10), "samples_per_sec": b_small * (n - 10) / (ms * 1e-3)}

    # ---- the drop-in with the UNEDITED trainer calls: torch.optim.Adam, per-tower BCELoss sum, regulariser through
    # autograd, graphs recorded lazily (nothing of INTEGRATION.md's opt-in table)
    plain = Trainer(pkg, wl, args, dev, world, "torch", "loss", "torch", "lazy")
    for i in range(max(args.warmup, 4)):
        plain.step(dev_x[i % n_batches], dev_y[i % n_batches], domains[i % n_batches])
    n = max(args.steps, 10)
    ms = timed(lambda i: plain.step(dev_x[i % n_batches], dev_y[i % n_batches], domains[i % n_batches]), 0, n)
    extra["unedited_trainer"] = {"ms_per_step": ms / n, "samples_per_sec": B * n / (ms * 1e-3),
                                 "what": "torch.optim.Adam + sum of BCELoss + get_regularization_loss through autograd, "
                                         "lazy graphs: what an unmodified run.py calls"}
    del plain
    torch.cuda.empty_cache()

    # ---- the HEMP regroup loop (run.py:614-661) on this model: per candidate generate_mask + load_model_state +
    # optimizer reset + 5 bagging steps with prun_single_mask + 5 no_grad scoring passes
    extra["regroup"] = regroup_leg(tr, wl, dev, dev_x, dev_y, n_domains=min(wl.n_domain, 30), candidates=10)

    # ---- batch feeding: the trainer's DataLoader over device tensors (run.py:301-306) against
    # data.DeviceBatchLoader, and the test loop's per-batch .cpu().numpy() (run.py:725-727) against data.EvalAccumulator
    extra["feeding"] = feeding_leg(tr, dev, dev_x, dev_y, domains)

    # ---- mixed-domain eval (BASELINE configs[3]): one 64K batch over the 355 Cloud-Theme-shaped domains
    extra["cloudtheme_mixed"] = mixed_leg(args, pkg, dev, kernel_ms)
    torch.cuda.empty_cache()

    # ---- torch-eager on the same GPU: the oracle port's ops on cuda (SURVEY 8d "same-box GPU bar")
    if not args.no_cpu_baseline:
        try:
            res, sec = oracle_rate(wl, B, 8, 3, args.active, args.seed, args.dropout, device=str(dev))
            extra["gpu_eager_baseline"] = {"value": res["value"], "unit": "samples/s", "ms_per_step": sec * 1e3,
                                           "sample": res["sample"]}
        except torch.cuda.OutOfMemoryError as e:          # pragma: no cover
            extra["gpu_eager_baseline"] = {"unavailable": str(e)[:120]}
        torch.cuda.empty_cache()
    return extra


def hei_leg(model, B, dev, kernel_ms, peaks, args):
    """aread_hei_layer_fwd / _bwd launches alone (the main kernel plus its finalise / reduce kernels), timed on their
    stream back to back.  Algorithmic bytes: input + output once forward; z + d_out + input + d_in once backward.  The
    tensors of one launch (25-100 MB) fit the 126 MB L2, as they do inside a step, so the rate is an L2-assisted one;
    tools/bench_hei.py times the same calls with the L2 flushed and with AREAD_HEI_TC=0 for the CUDA-core kernels."""
    ho = importlib.import_module(PKG + ".hei_ops")
    out = {"paths": "tcgen05 = csrc/hei_tc.cu, cuda_cores = csrc/hei.cu (aread_hei_set_path); same entry points, same bytes",
           "rows": B, "layers": []}
    gen = torch.Generator(device=dev).manual_seed(0)
    for l, dims in enumerate(TOWER_DIMS):
        G = N_TOWER[l]
        k = EXPERT_DIMS[-1] if l == 0 else TOWER_DIMS[l - 1][-1]
        for j, n in enumerate(dims):
            rnd = lambda *s: torch.randn(*s, device=dev, generator=gen)
            zp, w, b = rnd(B, G * k), 0.3 * rnd(G, n, k), rnd(G, n)
            gamma, beta = 1 + 0.1 * rnd(G * n), 0.1 * rnd(G * n)
            rm, rv = torch.zeros(G * n, device=dev), torch.ones(G * n, device=dev)
            saved = None
            if j > 0:
                saved = torch.stack([zp.mean(0), 1 / zp.std(0), 1 / zp.std(0), -zp.mean(0) / zp.std(0)]).contiguous()
            d_out = rnd(B, G * n)
            fb, bb = 4 * B * G * (k + n), 4 * B * G * (2 * n + 2 * k)
            rec = {"level": l, "towers": G, "k": k, "n": n, "input_is_preactivation": j > 0}
            for name, flag in (("tcgen05", 1), ("cuda_cores", 0)):
                ho.set_path(flag, flag)
                z, st = ho.layer_fwd(zp, saved, 7, w, b, gamma, beta, rm, rv, G, k, n, True, False, args.dropout, 11)
                coef, _ = ho.bn_bwd_coef(z, d_out, st, False, args.dropout, 11, 9)
                f_ms = kernel_ms(lambda i: ho.layer_fwd(zp, saved, 7, w, b, gamma, beta, rm, rv, G, k, n, True, False,
                                                        args.dropout, 11), 3, 20)
                b_ms = kernel_ms(lambda i: ho.layer_bwd(z, d_out, st, coef, args.dropout, 9, 11, False, zp, saved, 7, w,
                                                        G, k, n), 3, 20)
                rec[name] = {"fwd_us": f_ms * 1e3, "fwd_frac_of_hbm": fb / (f_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "bwd_us": b_ms * 1e3, "bwd_frac_of_hbm": bb / (b_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
            ho.set_path(-1, -1)
            out["layers"].append(rec)
            k = n
    return out


def feeding_leg(tr, dev, dev_x, dev_y, domains, bs=1024, n_batches=64):
    from torch.utils.data import DataLoader, TensorDataset
    data = importlib.import_module(PKG + ".data")
    X, y = dev_x[0][:bs * n_batches], dev_y[0][:bs * n_batches]
    out = {"batch_size": bs, "batches": n_batches}

    def drain(loader):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for xb, yb in loader:
            pass
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n_batches * 1e3

    drain(data.DeviceBatchLoader(X, y, batch_size=bs))
    out["torch_dataloader_ms_per_batch"] = drain(DataLoader(TensorDataset(X, y), batch_size=bs, shuffle=True))
    out["device_batch_loader_ms_per_batch"] = drain(data.DeviceBatchLoader(X, y, batch_size=bs))
    model = tr.model.eval()
    d = domains[0]

    def eval_loop(accumulate):
        acc = data.EvalAccumulator()
        host = []
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with torch.no_grad():
            for s0 in range(0, bs * n_batches, bs):
                xb, yb = X[s0:s0 + bs], y[s0:s0 + bs]
                pred = model(xb, mode="domain_with_mask", domain_i=d)
                if accumulate:
                    acc.add(yb, pred, xb[:, model.domain_idx])
                else:
                    host.append((yb.squeeze().cpu().numpy(), pred.squeeze().cpu().numpy(),
                                 xb[:, model.domain_idx].cpu().numpy()))
        if accumulate:
            acc.result()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n_batches * 1e3
    eval_loop(True)
    out["eval_per_batch_host_copies_ms_per_batch"] = eval_loop(False)
    out["eval_accumulator_ms_per_batch"] = eval_loop(True)
    tr.model.train()
    return out


def mixed_leg(args, pkg, dev, kernel_ms, B=65536):
    """AREAD.forward_mixed (eval, every row under its own domain's mask) on a Cloud-Theme-shaped model with 355 domains:
    domain-sorted against shuffled rows at three mask densities -- sorted rows let the HEI kernel skip the (32-row
    tile, tower) pairs the masks prune -- and the reference's way of evaluating the same rows, one call per domain."""
    wl = importlib.import_module(PKG + ".workloads").WORKLOADS["cloudtheme"]()
    mixed_ops = importlib.import_module(PKG + ".mixed_ops")
    model = pkg.AREAD(np.asarray(wl.one_hot_field_dims), wl.embed_dim, wl.multi_hot_dict, n_tower=N_TOWER,
                      n_domain=wl.n_domain, base_model="mmoe", expert_dims=EXPERT_DIMS, tower_dims=TOWER_DIMS,
                      domain_idx=wl.domain_idx, device=dev, dropout=args.dropout, config=make_config(wl)).to(dev).eval()
    model.reset_for_mask_update()
    xs, _ = wl.batch_mixed(B, args.seed, sort=True)
    xu, _ = wl.batch_mixed(B, args.seed, sort=False)
    xs, xu = torch.from_numpy(xs).to(dev), torch.from_numpy(xu).to(dev)
    seg = np.bincount(xs[:, wl.domain_idx].cpu().numpy(), minlength=wl.n_domain)
    out = {"batch": B, "n_domain": wl.n_domain, "domains_present": int((seg > 0).sum()),
           "segment_rows_median": float(np.median(seg[seg > 0])), "segment_rows_min": int(seg[seg > 0].min()),
           "vocabularies": "assumed (dataset not bundled): 0.72 M users, 1.36 M items, 1,000 leaf / 100 L1 categories",
           "mode": "eval(), no_grad, AREAD.forward_mixed (= 'domain_with_mask' per row)", "by_active_percent": {}}
    for active in (0.7, 0.3, 0.1):
        np.random.seed(args.seed + int(active * 100))
        for d in range(wl.n_domain):
            model.domain_mask[d] = model.generate_mask("rand", d, init_active_percent=active)
        towers = np.mean([[len(a) for a in model.mask_info(model.domain_mask[d]).active_idx] for d in range(wl.n_domain)],
                         axis=0)
        ms_sorted = kernel_ms(lambda i: model.forward_mixed(xs), 3, 10)
        ms_shuffled = kernel_ms(lambda i: model.forward_mixed(xu), 3, 10)

        def hei_only(x):        # the HEI kernel alone (the trunk in front of it does not depend on the masks)
            ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
            ts = []
            for _ in range(8):
                mixed_ops.forward_mixed_eval(model, x, hei_events=ev)
                torch.cuda.synchronize(dev)
                ts.append(ev[0].elapsed_time(ev[1]))
            return float(np.median(ts[2:]))
        hei_sorted, hei_shuffled = hei_only(xs), hei_only(xu)
        rec = {"mean_active_towers_per_level": [round(float(t), 2) for t in towers],
               "sorted_ms": ms_sorted, "sorted_samples_per_sec": B / (ms_sorted * 1e-3), "shuffled_ms": ms_shuffled,
               "hei_kernel_ms_sorted": hei_sorted, "hei_kernel_ms_shuffled": hei_shuffled,
               "hei_tile_skip_speedup_sorted_vs_shuffled": hei_shuffled / hei_sorted}
        if active == 0.3:       # the reference's evaluation of the same rows: one call per domain (run.py:719-727)
            bounds = np.concatenate(([0], np.cumsum(seg)))
            parts = [(d, xs[bounds[d]:bounds[d + 1]]) for d in range(wl.n_domain) if seg[d] > 0]

            def per_domain(i):
                with torch.no_grad():
                    for d, xd in parts:
                        model(xd, mode="domain_with_mask", domain_i=d)
            fused = importlib.import_module(PKG + ".fused")
            keep, fused.USE_GRAPHS = fused.USE_GRAPHS, False      # 300+ one-off shapes: nothing worth recording
            try:
                ms_loop = kernel_ms(per_domain, 1, 2)
            finally:
                fused.USE_GRAPHS = keep
            rec["per_domain_calls_ms"] = ms_loop
            rec["speedup_vs_per_domain_calls"] = ms_loop / ms_sorted
        out["by_active_percent"][str(active)] = rec
    del model
    return out


def regroup_leg(tr, wl, dev, dev_x, dev_y, n_domains, candidates, update_steps=5, eval_steps=5):
    model, opt = tr.model, tr.opt
    optim = importlib.import_module(PKG + ".optim")
    fast = optim.FusedAdam(model.parameters(), lr=LR, betas=(0.9, 0.99), eps=1e-8, weight_decay=WD)
    model.fold_regularization_into(fast)
    model.fold_regularization_into(opt)
    x, y = dev_x[0], dev_y[0]
    crit = torch.nn.BCELoss()
    # gate statistics for generate_mask('mask_max_gate')
    model.reset_for_mask_update()
    for d in range(n_domains):
        model(x, mode="wo_mask", domain_i=d, memory_gate_value=True)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    model.save_model_state()
    n_steps = 0
    for d in range(n_domains):
        for z in range(candidates):
            tmp_mask = model.generate_mask(generate_mode="mask_max_gate", d=d, init_active_percent=0.6,
                                           random_modify_sigma=0.2)
            model.load_model_state()
            fast.reset()
            for _ in range(update_steps):
                preds = model(x, mode="domain_mask_bagging", current_mask=tmp_mask, tmp_memory_gate_value=True)
                loss = model.bagging_loss(preds, y) + model.get_regularization_loss(device=dev)
                model.zero_grad()
                loss.backward()
                fast.step()
                tmp_mask = model.prun_single_mask(d, tmp_mask, prun_ratio=0.05)
                n_steps += 1
            model.candidate_domain_mask[d].append(tmp_mask)
            with torch.no_grad():
                for _ in range(eval_steps):
                    pred = model(x, mode="domain_with_mask", current_mask=tmp_mask)
                    loss = crit(pred.squeeze(), y.squeeze().float()) + model.get_regularization_loss(device=dev)
                    model.add_eval_loss(loss.mean().item(), d=d, mask_z=z)
    for d in range(n_domains, wl.n_domain):            # domains not searched keep their mask
        model.candidate_domain_mask[d].append(model.domain_mask[d])
        model.add_eval_loss(0.0, d=d, mask_z=0)
    model.update_all_mask(regroup_times=1)
    model.reset_for_mask_update()
    model.load_model_state()
    torch.cuda.synchronize(dev)
    sec = time.perf_counter() - t0
    return {"seconds": sec, "domains": n_domains, "candidates_per_domain": candidates,
            "train_steps": n_steps, "scoring_passes": n_domains * candidates * eval_steps,
            "table_restores": n_domains * candidates + 1, "batch": int(x.shape[0]),
            "ms_per_candidate": sec * 1e3 / (n_domains * candidates),
            "what": "run.py:614-661 with FusedAdam.reset() for optimizer_fast and the one-launch save/load_model_state"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    wl = importlib.import_module(PKG + ".workloads").WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
