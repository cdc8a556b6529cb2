"""Per-source-line stall samples and executed instructions from `ncu --page source --print-source cuda,sass --csv`.
usage: python tools/ncu_lines.py <csv> [top_n]"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, cur_line, cur_src = "", "", ""
    agg = defaultdict(lambda: [0, 0, ""])
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            si, ie = r.index("# Samples"), r.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= max(si, ie):
            continue
        if r[0]:
            cur_line, cur_src = r[0], r[1]
        if r[2] in ("-", "...", ""):
            continue
        key = (cur_file, cur_line)
        agg[key][0] += int(r[si]) if r[si].isdigit() else 0
        agg[key][1] += int(r[ie]) if r[ie].isdigit() else 0
        agg[key][2] = cur_src
    tot_s = sum(v[0] for v in agg.values())
    tot_i = sum(v[1] for v in agg.values())
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / max(tot_s, 1):5.1f}% smp {100 * v[1] / max(tot_i, 1):5.1f}% ins  {f}:{l}  {v[2].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
