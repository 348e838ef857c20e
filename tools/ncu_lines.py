#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).

usage: ncu_lines.py REPORT KERNEL_REGEX [TOP]
Prints stall samples and executed warp instructions per (file, line), sorted by samples, plus totals per file."""
import collections, csv, subprocess, sys, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
cur = None
hdr = None
lines = {}
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur = os.path.basename(row[1]); continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None or cur is None or row[0] == "":
        continue
    try:
        ln = int(row[0])
        smp = int(row[hdr["# Samples"]]); ins = int(row[hdr["Instructions Executed"]])
    except (ValueError, KeyError):
        continue
    stalls = {k: int(row[i]) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k and row[i].isdigit()}
    lines[(cur, ln)] = (smp, ins, row[1].strip()[:90], stalls)
tot_s = sum(v[0] for v in lines.values()) or 1
tot_i = sum(v[1] for v in lines.values()) or 1
print(f"total samples {tot_s}  total warp instructions {tot_i}")
byfile = collections.defaultdict(lambda: [0, 0])
for (f, ln), v in lines.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} samples {100*s/tot_s:5.1f}%  instr {100*i/tot_i:5.1f}%")
print()
for (f, ln), (s, i, src, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    top_st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{f}:{ln:<5d} smp {100*s/tot_s:5.1f}% ins {100*i/tot_i:5.1f}%  [{top_st}]  {src}")
