"""Stall-reason breakdown of the top source lines (input: ncu --page source --print-source cuda,sass --csv)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = None
cur = ("", "")
agg = defaultdict(lambda: defaultdict(int))
tot = defaultdict(int)
f = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0]:
        cur = (f, r[0])
    if r[2] in ("-", "...", ""):
        continue
    for i in stall:
        if r[i].isdigit():
            agg[cur][hdr[i][6:]] += int(r[i])
            tot[hdr[i][6:]] += int(r[i])
allsum = sum(tot.values())
print("all:", [(k, round(100 * v / allsum, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]])
for key, d in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(d.values())
    print(f"{100 * s / allsum:5.1f}%  {key[0]}:{key[1]}  ", [(k, v) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:4]])
