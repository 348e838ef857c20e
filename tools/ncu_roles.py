#!/usr/bin/env python3
"""Stall samples of the two-warps-per-cell backward kernel split by warp role (solver / accumulator), from the source page
of an .ncu-rep: usage ncu_roles.py REPORT [KERNEL_REGEX]"""
import collections, csv, os, subprocess, sys
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_ros_adj_pair"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
HERE = os.path.dirname(os.path.abspath(__file__))
pair = open(os.path.join(HERE, "..", "fatode_b200", "csrc", "ros_cell_adj_pair.cuh")).read().splitlines()
split = next(i for i, l in enumerate(pair, 1) if "void adjp_accumulate" in l)          # first line of the accumulator's code
kern_line = next(i for i, l in enumerate(pair, 1) if "k_ros_adj_pair(RosParams" in l)
dev = open(os.path.join(HERE, "..", "fatode_b200", "csrc", "gen", "cbm4_dev_gen.h")).read().splitlines()
jt0 = next(i for i, l in enumerate(dev, 1) if "void jt_vec(" in l)
mt0 = next(i for i, l in enumerate(dev, 1) if "void mt_rows(" in l)


def role(f, ln):
    if f == "cbm4_cell_gen.h":
        return "solver"
    if f == "cbm4_dev_gen.h":
        return "solver" if jt0 <= ln < mt0 else "accumulator"
    if f == "ros_cell_adj_pair.cuh":
        return "shared" if ln >= kern_line - 3 else ("solver" if ln < split - 3 else "accumulator")
    if f == "ros_cell_adj.cuh":
        return "mbar/tma helpers"
    return "other (" + f + ")"


cur = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0] == "File Path": cur = os.path.basename(row[1]); continue
    if row[0] == "Line No": hdr = {h: i for i, h in enumerate(row)}; continue
    if hdr is None or cur is None or row[0] == "": continue
    try: ln = int(row[0]); smp = int(row[hdr["# Samples"]]); ins = int(row[hdr["Instructions Executed"]])
    except (ValueError, KeyError): continue
    a = agg[role(cur, ln)]
    a["samples"] += smp; a["instructions"] += ins
    for k, i in hdr.items():
        if k.startswith("stall_") and "Not Issued" not in k and row[i].isdigit(): a[k[6:]] += int(row[i])
tot = sum(a["samples"] for a in agg.values()) or 1
for r, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = ", ".join(f"{k} {100*v/max(a['samples'],1):.0f}%" for k, v in a.most_common(12) if k not in ("samples", "instructions"))
    print(f"{r:28s} samples {100*a['samples']/tot:5.1f}%  warp instructions {a['instructions']:>12d}  cycles/instr {a['samples'] and 1:d}\n    {st}")
