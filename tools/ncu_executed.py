#!/usr/bin/env python3
"""profiles/executed_counts.json from an ncu capture of tools/prof_case.py (all launches of the three headline kernels).

    python tools/ncu_executed.py <capture.ncu-rep> <prof_case_counts.json> <library.so> [capture-name]

Sums smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on and dram bytes over every captured launch of each kernel and
divides by the accepted steps / attempts the same run reports (prof_case.py writes them next to its output).  FMA counts 2.
bench.py reports executed-flop fractions only when the SHA-256 in this file matches the library it runs."""
import csv, hashlib, json, subprocess, sys

rep, counts_json, lib = sys.argv[1], sys.argv[2], sys.argv[3]
name = sys.argv[4] if len(sys.argv) > 4 else rep
cnt = json.load(open(counts_json))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = {h: i for i, h in enumerate(rows[0])}
units = rows[1]


def num(r, key):
    v = float(r[hdr[key]].replace(",", ""))
    u = units[hdr[key]].lower()
    if "byte" in u:
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return v


agg = {}
for r in rows[2:]:
    k = r[hdr["Kernel Name"]]
    key = "forward" if "k_ros_fwd_cell" in k else "expand" if "k_expand_tape" in k else "adjoint" if "k_ros_adj_" in k else None
    if key is None:
        continue
    a = agg.setdefault(key, dict(kernel=k.split("(")[0], launches=0, dfma=0.0, dmul=0.0, dadd=0.0, dram=0.0, ns=0.0, inst=0.0))
    a["launches"] += 1
    a["dfma"] += num(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
    a["dmul"] += num(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
    a["dadd"] += num(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
    a["dram"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    a["inst"] += num(r, "smsp__inst_executed.sum")
out = {"capture": name, "library_sha256": hashlib.sha256(open(lib, "rb").read()).hexdigest(), "case": cnt}
for key, a in agg.items():
    flops = 2.0 * a["dfma"] + a["dmul"] + a["dadd"]
    per = cnt["attempts"] if key == "forward" else cnt["accepted_steps"]
    unit = "attempt" if key == "forward" else "accepted_step"
    out[key] = {"kernel": a["kernel"], "launches": a["launches"], f"fp64_flops_per_{unit}": flops / per,
                f"dfma_per_{unit}": a["dfma"] / per, f"dmul_dadd_per_{unit}": (a["dmul"] + a["dadd"]) / per,
                f"dram_bytes_per_{unit}": a["dram"] / per, f"warp_instructions_per_{unit}": a["inst"] / per}
json.dump(out, sys.stdout, indent=1)
print()
