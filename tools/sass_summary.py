#!/usr/bin/env python3
"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that show what the shipped library executes
(tensor memory LDTM/STTM, TMEM allocation UTCATOMSWS, bulk TMA UBLKCP, mbarrier SYNCS, FP64 DFMA/DADD/DMUL, shared/global/local
memory traffic) from one `cuobjdump -sass` of fatode_b200/lib/libfatode_b200.so.   usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fatode_b200", "lib", "libfatode_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
KEYS = ["LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "UTC"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = re.sub(r"^void ", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        per[cur]["total"] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
                break
print(f"library sha256 {hashlib.sha256(open(SO, 'rb').read()).hexdigest()}  (cuobjdump -sass fatode_b200/lib/libfatode_b200.so; static instruction counts)")
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>6s}" for k in KEYS[:-1]) + "  UTC*MMA")
tot = collections.Counter()
for k, c in per.items():
    if c["total"] < 50:
        continue
    print(f"{k[:70]:70s} {c['total']:7d} " + " ".join(f"{c[x]:6d}" for x in KEYS[:-1]) + f"  {c['UTC']:6d}")
    tot.update(c)
print(f"{'ALL KERNELS':70s} {tot['total']:7d} " + " ".join(f"{tot[x]:6d}" for x in KEYS[:-1]) + f"  {tot['UTC']:6d}")
