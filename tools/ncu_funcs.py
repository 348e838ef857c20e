#!/usr/bin/env python3
"""Aggregate an .ncu-rep source page by C++ function (of one header): usage ncu_funcs.py REPORT KERNEL_REGEX HEADER"""
import csv, os, re, subprocess, sys, bisect, collections
rep, kern, hdr = sys.argv[1], sys.argv[2], sys.argv[3]
starts, names = [], []
for i, l in enumerate(open(hdr), 1):
    m = re.match(r"\s*(?:CBM4_CELL_FN|CBM4_CELL_INL|template.*|__device__.*|static.*)?\s*(?:void|int|double|unsigned long long)\s+(\w+)\(", l)
    if m and ("CBM4_CELL" in l or "__device__" in l or "__global__" in l):
        starts.append(i); names.append(m.group(1))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
cur = None; h = None
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0] == "File Path": cur = os.path.basename(row[1]); continue
    if row[0] == "Line No": h = {k: i for i, k in enumerate(row)}; continue
    if h is None or cur is None or row[0] == "": continue
    try: ln = int(row[0]); s = int(row[h["# Samples"]]); n = int(row[h["Instructions Executed"]])
    except Exception: continue
    tot[0] += s; tot[1] += n
    if cur == os.path.basename(hdr):
        k = bisect.bisect_right(starts, ln) - 1
        key = names[k] if k >= 0 else "?"
    else:
        key = "[" + cur + "]"
    agg[key][0] += s; agg[key][1] += n
for k, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} samples {100*s/max(tot[0],1):5.1f}%  instr {100*n/max(tot[1],1):5.1f}%  ({n} instr, ipc-ish {n/max(s,1):.0f} instr/sample)")
