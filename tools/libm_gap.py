#!/usr/bin/env python3
"""How much the Rosenbrock step sequence of CBM-IV depends on the last bit of COS / Err**(1/4).
Runs the CPU oracle over N perturbed cells of the headline problem with (a) the correctly rounded kernels (default mode) and
(b) the host libm (ORACLE_LIBM mode) and reports the fraction of cells with identical ISTATUS and the state / Lambda
differences.  CPU only (about 2 x N / 9 seconds on 8 threads).   python tools/libm_gap.py [N] > profiles/r2_libm_gap.json"""
import importlib.util, json, os, platform, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
spec = importlib.util.spec_from_file_location("inp", os.path.join(ROOT, "fatode_b200", "inputs.py"))
inp = importlib.util.module_from_spec(spec); spec.loader.exec_module(inp)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F01 = float(np.float32(0.1)); T0 = 43200.0; T1 = T0 + 72 * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y0 = inp.perturbed_states(O.initial_state("cbm4"), n)
res = {}
for mode in (0, 1):
    O.lib().oracle_set_libm(mode)
    res[mode] = O.ros_adj("cbm4", y0, 32, T0, T1, atol, rtol, ic, with_mu=False)
O.lib().oracle_set_libm(0)
a, b = res[0], res[1]
same = np.all(a["istatus"] == b["istatus"], axis=1)
dy = np.max(np.abs(a["y"] - b["y"]), axis=1) / np.max(np.abs(b["y"]), axis=1)
cs = np.max(np.abs(b["lambda_"]), axis=2)
dl = np.max(np.abs(a["lambda_"] - b["lambda_"]), axis=2) / np.maximum(cs, 1e-10 * cs.max(axis=1, keepdims=True))
dn = a["istatus"][:, 3] - b["istatus"][:, 3]
out = {"cells": n, "host_libm": platform.libc_ver(), "identical_istatus_fraction": float(same.mean()),
       "nacc_difference": {"min": int(dn.min()), "max": int(dn.max()), "mean_abs": float(np.abs(dn).mean())},
       "state_rel_diff": {"identical_sequence_max": float(dy[same].max()) if same.any() else None,
                          "different_sequence_max": float(dy[~same].max()) if (~same).any() else None, "median": float(np.median(dy))},
       "lambda_col_rel_diff": {"identical_sequence_max": float(dl[same].max()) if same.any() else None,
                               "different_sequence_max": float(dl[~same].max()) if (~same).any() else None}}
print(json.dumps(out, indent=1))
