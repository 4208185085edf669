#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per kernel, executed warp instructions by opcode
class and the top stall reasons.  Usage: summarize_src.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
kern = None; hdr = None
stats = {}
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        kern = r[1][:60]; stats[kern] = dict(ops=collections.Counter(), stalls=collections.Counter(), samples=0, inst=0); hdr = None; continue
    if r[0] == "Address":
        hdr = r; continue
    if hdr is None or kern is None: continue
    d = dict(zip(hdr, r))
    sass = d["Source"].strip()
    op = sass.split()[0] if sass else "?"
    if op.startswith("@"): op = sass.split()[1]
    op = op.split(".")[0]
    n = int(d["Instructions Executed"] or 0)
    st = stats[kern]
    st["ops"][op] += n; st["inst"] += n
    st["samples"] += int(d["# Samples"] or 0)
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
            st["stalls"][k] += int(v)
for k, st in stats.items():
    print("==", k, "warp-inst", st["inst"], "samples", st["samples"])
    print("  ops:", ", ".join("%s %.1f%%" % (o, 100.0 * c / max(1, st["inst"])) for o, c in st["ops"].most_common(18)))
    tot = sum(st["stalls"].values())
    print("  stalls:", ", ".join("%s %.1f%%" % (o, 100.0 * c / max(1, tot)) for o, c in st["stalls"].most_common(8)))
