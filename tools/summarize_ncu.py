#!/usr/bin/env python
"""Summaries of the ncu captures for profiles/: per-kernel share of a launch list, key metrics of a --set full report."""
import collections, csv, re, subprocess, sys

def launches(path, first_step=None, last_step=None):
    """Per-kernel totals.  With first_step / last_step only the launches of those train steps are counted: a step ends
    with its adam_kernel launch, so step k spans (adam #k-1, adam #k] (1-based, in launch order)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; n = 0
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if first_step is not None:
        ends = [i for i, r in enumerate(rows) if "adam_kernel" in r["Kernel Name"]]
        lo = ends[first_step - 2] + 1 if first_step >= 2 else 0
        rows = rows[lo:ends[last_step - 1] + 1]
    for row in rows:
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = row["Kernel Name"]
        m = re.search(r"(\w+_kernel)", name)
        key = ("aread::" + m.group(1)) if ("aread" in name and m) else re.sub(r"[<(].*", "", name)[-60:]
        agg[key][0] += 1; agg[key][1] += v; tot += v; n += 1
    out = [f"# {path}: {n} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)",
           f"{'us':>10} {'count':>6} {'share':>7}  kernel"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        out.append(f"{t:10.1f} {c:6d} {100 * t / tot:6.1f}%  {k}")
    mine = sum(t for k, (c, t) in agg.items() if k.startswith("aread::"))
    out.append(f"# library kernels: {100 * mine / tot:.1f}% of the captured GPU time")
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]

def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# {path}"]
    for r in rows[2:]:
        out.append("kernel: " + r[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                out.append(f"    {w:70s} {r[idx[w]]} {units[idx[w]]}")
    return "\n".join(out)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2], *(int(v) for v in sys.argv[3:5])))
    else:
        print(full(sys.argv[2]))
