"""First contact of the per-thread path on the GPU: small batch, compare with the general path and the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fatode_b200 as F
import oracle_lib as O
from fatode_b200.inputs import perturbed_states
F01 = float(np.float32(0.1)); T0 = 43200.0
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T1 = T0 + hours * 3600.0
rtol, atol = np.full(32, F01 * 1e-4), np.full(32, F01 * 1e-6)
ic = np.zeros(20, dtype=np.int32); ic[2] = 5
y0 = perturbed_states(F.initial_state("cbm4"), nc)
ref = O.ros_adj("cbm4", y0, 32, T0, T1, atol, rtol, ic)
f = F.integrate("cbm4", T0, T1, y0, rtol, atol, icntrl_u=ic)
rf = O.ros_fwd("cbm4", y0, T0, T1, atol, rtol, ic)
s = F.last_stats()
print("FWD family: ierr", f.ierr[:4], "istatus equal", np.array_equal(f.istatus, rf["istatus"]), "y equal", np.array_equal(f.y, rf["y"]),
      "general", s.cells_general, "ms", s.ms_forward)
r = F.integrate_adj("cbm4", T0, T1, y0, 32, rtol, atol, icntrl_u=ic)
s = F.last_stats()
print("ADJ: ierr", r.ierr[:4], "istatus equal", np.array_equal(r.istatus, ref["istatus"]), "y equal", np.array_equal(r.y, ref["y"]))
print("istatus gpu", r.istatus[0, :8], "ref", ref["istatus"][0, :8])
sc = np.max(np.abs(ref["lambda_"]), axis=2); er = np.max(np.abs(r.lambda_ - ref["lambda_"]), axis=2)
print("lambda max err/colscale (cols > 1e-10 of max):", np.max((er / np.maximum(sc, 1e-10 * sc.max(axis=1, keepdims=True)))))
ps = np.max(np.abs(ref["mu"]), axis=1, keepdims=True)
print("mu max err/rowscale:", np.max(np.abs(r.mu - ref["mu"]) / np.maximum(ps, 1e-300)))
print(f"stats: waves {s.waves} launches {s.launches} fwd {s.ms_forward:.2f} adj {s.ms_adjoint:.2f} general {s.cells_general} deferred {s.cells_deferred} steps {s.accepted_steps}")
