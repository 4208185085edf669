#!/usr/bin/env python
"""Per-CUDA-line executed warp instructions and stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
Usage: summarize_lines.py file.csv [kernel-substring] [min_pct]   (lines in source order)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
kern = fname = None
data = {}
for r in rows:
    if not r: continue
    if r[0] == "Function Name": kern = r[1].split("(")[0].replace("void tbk::", ""); continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No" or not r[0].isdigit(): continue
    try:
        samples, inst = int(r[6]), int(r[7])
    except ValueError:
        continue
    if inst and pat in (kern or ""):
        key = (kern, fname, int(r[0]))
        old = data.get(key, (0, 0, ""))
        data[key] = (old[0] + inst, old[1] + samples, r[1].strip()[:90])
tot = {}
for (k, f, l), v in data.items():
    t = tot.setdefault(k, [0, 0]); t[0] += v[0]; t[1] += v[1]
last = None
for key in sorted(data):
    k, f, l = key; v = data[key]
    if k != last:
        print("==", k, "warp-inst", tot[k][0], "samples", tot[k][1]); last = k
    pct = 100.0 * v[0] / tot[k][0]
    if pct >= minpct:
        print("  %5.1f%% inst %5.1f%% smp  %s:%d  %s" % (pct, 100.0 * v[1] / max(tot[k][1], 1), f, l, v[2]))
